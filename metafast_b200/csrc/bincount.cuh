// bincount.cuh -- bin-local counting: the default counting kernels of MFKC_VARIANT_HASH for k <= 31.
//
// The extraction kernel files every super-k-mer record under the BIN of its minimizer (bin = a fixed
// pseudo-random function of the canonical minimizer, hence of the canonical k-mer alone: every instance of a
// k-mer lands in the same bin).  Bins are sized so that the distinct k-mers of one bin fit a hash table in
// SHARED memory, so the count of a whole sample is
//     for every bin (one CTA each, persistent CTAs fetch bins from an atomic counter):
//         stream the bin's records into shared memory with TMA bulk copies (cp.async.bulk + mbarrier,
//         one 512-byte copy per warp in flight), re-expand them into canonical k-mers, upsert them into
//         the shared-memory table (LDS + ATOMS, no L2 round trip per k-mer), then sweep the table once:
//         histogram, distinct count, compaction of the entries with count > threshold.
// HBM sees the records once (2.7 B per k-mer instance) and the selected (key, count) pairs once: there is
// no global table to clear, sweep or scan, and memory is 2.7 B per instance instead of 40-55 B per distinct
// k-mer.  Replaces Long2ShortHashMap.addAndBound ([itmo]/structures/map/Long2ShortHashMap.java:119-157) and
// the iteration + filter of IOUtils.printKmers (src/io/IOUtils.java:57-66) with identical results.
//
// Exactness under skew:
//   * the records of a bin that did not fit its staging segment continue in 32-record chunks of an overflow pool
//     (found through a tag table, ovf_chunk_of) and are streamed in like the segment's own batches;
//   * a (sub-)pass that fills the shared-memory table aborts without output and is split by key hash into
//     2..32 sub-ranges that are counted one after the other (the records are re-read from L2); a sub-range
//     that still does not fit at 32 parts is HEAVY;
//   * heavy (bin, sub-range) entries are counted by drain_heavy_kernel into the global table (the legacy path of
//     kernels.cuh) and emitted by table_scan_kernel into the same output arrays.  A k-mer belongs to exactly one (bin, sub-range), so nothing is counted twice.
#pragma once
#include "kernels.cuh"

namespace mfkc {

constexpr int BC_LOG2S = 13;               // slots of the shared-memory table (8192 x (8 B key + 4 B count) = 96 KiB)
constexpr int BC_THREADS = 512;            // 16 warps; 2 CTAs per SM
constexpr int BC_HIST = 256;               // histogram bins kept in shared memory (higher counts: global atomics)
constexpr int BC_MAX_SRC = P2P_MAX_PEERS;  // staging buffers one bin is gathered from (1 on one GPU, G with peer memory)
constexpr uint32_t BC_MAXP = 32;           // most sub-ranges a bin is split into before it is declared heavy
constexpr uint32_t BC_MAX_PROBE = 256;     // probes after which an upsert gives up (the pass aborts and is split)

struct BinSrc {
    const uint4 *recs[BC_MAX_SRC];             // source s: segment (seg0 + bin) of seg_cap records
    const unsigned int *cursor[BC_MAX_SRC];    // records appended per segment (> seg_cap: the surplus is in chunks of the overflow pool)
    const uint4 *ovf[BC_MAX_SRC];              // overflow pool of source s: ovf_chunks chunks of 32 records ...
    unsigned long long *ovf_tags[BC_MAX_SRC];  // ... and its tag table (ovf_chunk_of)
    uint32_t ovf_chunks;
    uint64_t seg_cap;
    uint32_t n_src;
    uint32_t seg0;                             // first segment of this shard in every source (shard * n_bins)
    uint32_t rot;                              // source visited first (own GPU), spreads the NVLink load
};
struct HeavyEnt { uint32_t bin; uint16_t p, P; };             // keys of `bin` with (bc_hash & (P-1)) == p
struct BinCtl {
    unsigned int next_bin, n_heavy, heavy_overflow, n_split;     // per count pass (zeroed up to ovf_cursor before every pass)
    unsigned long long heavy_recs, total_recs;
    unsigned int ovf_cursor, pad0;                                // per sample: fill of the overflow list
};
struct BinCountArgs {
    BinSrc src;
    uint32_t n_bins;
    int k;
    uint32_t thr;                              // output entries with count > thr
    uint32_t limit;                            // claimed slots at which a pass aborts (< 2^BC_LOG2S)
    uint32_t max_recs;                         // bins with more records go straight to the heavy list
    uint32_t bins_per_cta;                     // a CTA retires after this many bins, so that CTAs of other streams get SMs
    int use_tma;
    unsigned long long *out_keys; uint16_t *out_counts; uint64_t out_cap;
    unsigned long long *hist;
    Counters *ctr;
    BinCtl *ctl;
    HeavyEnt *heavy; uint32_t heavy_cap;
};

// slot / sub-range hash of the shared-memory table: the top bits pick the slot, the low bits the sub-range
__host__ __device__ __forceinline__ uint32_t bc_hash(uint64_t key) {
    uint32_t h = ((uint32_t)key * 0x9E3779B1u) ^ ((uint32_t)(key >> 32) * 0x85EBCA6Bu);
    h ^= h >> 15;
    h *= 0xC2B2AE35u;
    return h ^ (h >> 13);
}

// ---- mbarrier / TMA bulk copy (PTX; SASS: SYNCS.*, UBLKCP) ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spins = 0; !ok; spins++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (spins > (1u << 26)) asm volatile("trap;");         // a copy that never lands must not hang the device
    }
}

// shared-memory table accesses by 32-bit shared-space address (the table base stays in one register; no generic-to-shared
// conversions in the probe loop)
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long cas_shared_u64(uint32_t a, unsigned long long cmp, unsigned long long val) {
    unsigned long long old;
    asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(a), "l"(cmp), "l"(val) : "memory");
    return old;
}
__device__ __forceinline__ void inc_shared_u32(uint32_t a) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); }

template <int LOG2S, int NT>
constexpr size_t bin_count_smem_bytes() {
    return ((size_t)12 << LOG2S) + (size_t)(NT / 32) * (512 + 2 * 16 * 4) + BC_HIST * 4 + (size_t)(NT / 32) * 8 + (NT / 32) * 4 + 64 * 4 + (2 * BC_MAX_SRC + 8) * 4 + 64;
}

template <int LOG2S, int NT>
__global__ void __launch_bounds__(NT, 2)
bin_count_kernel(const __grid_constant__ BinCountArgs a) {
    constexpr uint32_t S = 1u << LOG2S, NW = NT / 32, U = S / NT;
    constexpr uint32_t FULL = 0xffffffffu;
    extern __shared__ __align__(128) uint8_t bc_smem[];
    unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(bc_smem);            // S x 8
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_keys + S);                               // S x 4
    uint4 *s_ring = reinterpret_cast<uint4 *>(s_cnt + S);                                     // NW x 32 records (TMA destination)
    uint32_t *s_w = reinterpret_cast<uint32_t *>(s_ring + NW * 32);                           // NW x 16: record-start bits
    uint32_t *s_b = s_w + NW * 16;                                                            // NW x 16: records before word j
    uint32_t *s_hist = s_b + NW * 16;                                                         // BC_HIST
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_hist + BC_HIST);     // NW mbarriers
    uint32_t *s_wgood = reinterpret_cast<uint32_t *>(s_bar + NW);                             // NW
    uint32_t *s_stack = s_wgood + NW;                                                         // 64 (p | P << 16)
    uint32_t *s_srcstart = s_stack + 64;                                                      // BC_MAX_SRC + 1: first batch of every source
    uint32_t *s_srcn = s_srcstart + BC_MAX_SRC + 1;                                           // BC_MAX_SRC: records of every source (+ 7 pad)
    uint32_t *s_ctl = s_srcn + BC_MAX_SRC + 7;                                                // bin, sp, claimed, abort, done, next batch
    volatile uint32_t *vs_ctl = s_ctl;
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(s_ctl + 8);
    enum { C_BIN = 0, C_SP, C_CLAIMED, C_ABORT, C_DONE, C_NEXTB };

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t lane_le = lanemask_lt() | (1u << lane);
    const int rs = 64 - 2 * a.k;
    const uint32_t keys_a = smem_u32(s_keys), cnt_a = smem_u32(s_cnt);       // shared-space addresses of the table
    for (uint32_t i = tid; i < S; i += NT) { s_keys[i] = EMPTY_KEY; s_cnt[i] = 0; }
    for (uint32_t i = tid; i < BC_HIST; i += NT) s_hist[i] = 0;
    const uint32_t bar = smem_u32(&s_bar[warp]);
    const uint32_t ring = smem_u32(&s_ring[warp * 32]);
    if (lane == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint32_t parity = 0;
    unsigned long long occ_total = 0, recs_total = 0;

    for (uint32_t taken = 0; taken < a.bins_per_cta; taken++) {
        if (tid == 0) s_ctl[C_BIN] = atomicAdd(&a.ctl->next_bin, 1u);
        __syncthreads();
        const uint32_t bin = vs_ctl[C_BIN];
        if (bin >= a.n_bins) break;
        // records of this bin in every source (segment + overflow chunks)
        uint64_t n_recs = 0; const bool light = true;
        for (uint32_t s = 0; s < a.src.n_src; s++) n_recs += a.src.cursor[s][a.src.seg0 + bin];
        if (tid == 0) recs_total += n_recs;
        if (n_recs == 0 && light) { __syncthreads(); continue; }
        if (!light || n_recs > a.max_recs) {
            if (tid == 0) {
                const uint32_t i = atomicAdd(&a.ctl->n_heavy, 1u);
                if (i < a.heavy_cap) { HeavyEnt e; e.bin = bin; e.p = 0; e.P = 1; a.heavy[i] = e; } else a.ctl->heavy_overflow = 1u;
                atomicAdd(&a.ctl->heavy_recs, (unsigned long long)n_recs);
            }
            __syncthreads();
            continue;
        }
        if (tid == 0) { s_stack[0] = 0u | (1u << 16); s_ctl[C_SP] = 1; }
        __syncthreads();

        for (;;) {                                                   // sub-ranges of this bin, depth first
            const uint32_t sp = vs_ctl[C_SP];
            if (sp == 0) break;
            const uint32_t item = s_stack[sp - 1];
            __syncthreads();
            if (tid == 0) { s_ctl[C_SP] = sp - 1; s_ctl[C_CLAIMED] = 0; s_ctl[C_ABORT] = 0; s_ctl[C_DONE] = 0; }
            __syncthreads();
            const uint32_t p = item & 0xFFFFu, P = item >> 16;

            // ---- count pass: the warps take batches of 32 records from a shared counter (records hold 1..16 k-mers, so
            // batches differ in work); the batches of all sources form one index space
            if (tid == 0) {
                uint32_t run = 0;
                for (uint32_t sj = 0; sj < a.src.n_src; sj++) {
                    const uint32_t sidx = (sj + a.src.rot) % a.src.n_src;
                    const uint32_t n = a.src.cursor[sidx][a.src.seg0 + bin];
                    const uint32_t n_seg = n > a.src.seg_cap ? (uint32_t)a.src.seg_cap : n;
                    s_srcn[sj] = n; s_srcstart[sj] = run;
                    run += ((n_seg + 31) >> 5) + ((n - n_seg + 31) >> 5);      // batches of the segment, then one per overflow chunk
                }
                s_srcstart[a.src.n_src] = run;
                s_ctl[C_NEXTB] = 0;
            }
            __syncthreads();
            const uint32_t n_batches_total = s_srcstart[a.src.n_src];
            {
                // batch g -> (source, first record, records in the batch); lane 0 fetches, everybody gets the same answer
                uint32_t cnt = 0; const uint4 *src_ptr = nullptr;
                auto fetch = [&]() -> bool {
                    uint32_t g = 0;
                    if (lane == 0) g = atomicAdd(&s_ctl[C_NEXTB], 1u);
                    g = __shfl_sync(FULL, g, 0);
                    if (g >= n_batches_total) return false;
                    uint32_t sj = 0;
                    while (g >= s_srcstart[sj + 1]) sj++;
                    const uint32_t bb = g - s_srcstart[sj], n = s_srcn[sj];
                    const uint32_t sidx = (sj + a.src.rot) % a.src.n_src;
                    const uint32_t n_seg = n > a.src.seg_cap ? (uint32_t)a.src.seg_cap : n;
                    const uint32_t seg_batches = (n_seg + 31) >> 5;
                    if (bb < seg_batches) {
                        src_ptr = a.src.recs[sidx] + (uint64_t)(a.src.seg0 + bin) * a.src.seg_cap + (size_t)bb * 32;
                        cnt = min(32u, n_seg - bb * 32);
                    } else {                                         // overflow chunk j of this bin in this source
                        const uint32_t j = bb - seg_batches;
                        uint32_t ch = 0;
                        if (lane == 0) ch = ovf_chunk_of<false>(a.src.ovf_tags[sidx], a.src.ovf_chunks, a.src.seg0 + bin, j, nullptr);
                        ch = __shfl_sync(FULL, ch, 0);
                        src_ptr = a.src.ovf[sidx] + (size_t)(ch == OVF_NO_CHUNK ? 0u : ch) * 32;
                        cnt = ch == OVF_NO_CHUNK ? 0u : min(32u, n - n_seg - j * 32);      // (a missing chunk = dropped records: reported by mfkc_flush)
                    }
                    return true;
                };
                uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
                auto issue = [&]() {                                 // start moving the fetched batch: TMA bulk copy, or a plain load
                    if (a.use_tma) {
                        if (lane == 0) {
                            if (cnt) { mbar_expect_tx(bar, cnt * 16u); bulk_g2s(ring, src_ptr, cnt * 16u, bar); }
                            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");      // empty batch: complete the phase
                        }
                    } else if (lane < cnt) nxt = ld_nc_u128(src_ptr + lane);
                };
                bool have = fetch();
                if (have) issue();
                while (have) {
                    const uint32_t ab = __shfl_sync(FULL, lane == 0 ? vs_ctl[C_ABORT] : 0u, 0);
                    if (ab) break;
                    const uint32_t cur_cnt = cnt;
                    uint4 r;
                    if (a.use_tma) {
                        mbar_wait(bar, parity); parity ^= 1u;
                        r = lane < cur_cnt ? s_ring[warp * 32 + lane] : make_uint4(0u, 0u, 0u, 0u);
                        __syncwarp();
                    } else r = nxt;
                    have = fetch();                                  // the ring is free again: the next copy flies during the expansion
                    if (have) issue();
                    const uint32_t len = lane < cur_cnt ? (r.z & 15u) + 1u : 0u;
                    uint32_t incl = len;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += v; }
                    const uint32_t excl = incl - len;
                    const uint32_t total = __shfl_sync(FULL, incl, 31);          // <= 512 instances
                    // instance t belongs to the record whose start bit is the last one at or before t
                    if (lane < 16) s_w[warp * 16 + lane] = 0u;
                    __syncwarp();
                    if (len) atomicOr(&s_w[warp * 16 + (excl >> 5)], 1u << (excl & 31));
                    __syncwarp();
                    {
                        const uint32_t pc = lane < 16 ? __popc(s_w[warp * 16 + lane]) : 0u;
                        uint32_t ip = pc;
#pragma unroll
                        for (int o = 1; o < 16; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, ip, o); if (lane >= (uint32_t)o) ip += v; }
                        if (lane < 16) s_b[warp * 16 + lane] = ip - pc;
                    }
                    __syncwarp();
                    uint32_t wclaimed = 0;
                    for (uint32_t t0 = 0, j = 0; t0 < total; t0 += 32, j++) {
                        const uint32_t t = t0 + lane;
                        const uint32_t w = s_w[warp * 16 + j], bs = s_b[warp * 16 + j];
                        const uint32_t lo = (bs + __popc(w & lane_le) - 1u) & 31u;
                        const uint32_t qx = __shfl_sync(FULL, r.x, lo), qy = __shfl_sync(FULL, r.y, lo), qz = __shfl_sync(FULL, r.z, lo);
                        const uint32_t qs = __shfl_sync(FULL, excl, lo);
                        bool claimed_now = false;
                        if (t < total) {
                            const uint32_t off = 2u * (t - qs);
                            const uint32_t w2 = qz & ~15u;
                            const uint32_t h32 = __funnelshift_l(qy, qx, off);
                            const uint32_t l32 = __funnelshift_l(w2, qy, off);
                            const uint64_t fw = (((uint64_t)h32 << 32) | l32) >> rs;
                            const uint64_t rc = revcomp64(fw, a.k);
                            const unsigned long long key = fw < rc ? fw : rc;
                            const uint32_t h = bc_hash(key);
                            if ((h & (P - 1u)) == p) {
                                uint32_t slot = h >> (32 - LOG2S);
                                uint32_t probe = 0;
                                for (;; probe++) {
                                    const unsigned long long cur = lds_u64(keys_a + slot * 8u);
                                    if (cur == key) break;
                                    if (cur == EMPTY_KEY) {
                                        const unsigned long long prev = cas_shared_u64(keys_a + slot * 8u, EMPTY_KEY, key);
                                        if (prev == EMPTY_KEY) { claimed_now = true; break; }
                                        if (prev == key) break;
                                    }
                                    if (probe >= BC_MAX_PROBE) { s_ctl[C_ABORT] = 1u; slot = S; break; }      // pass is void; it will be split
                                    slot = (slot + 1u) & (S - 1u);
                                }
                                if (slot < S) inc_shared_u32(cnt_a + slot * 4u);               // one increment site for every exit of the probe loop
                            }
                        }
                        wclaimed += __popc(__ballot_sync(FULL, claimed_now));
                    }
                    if (lane == 0) {
                        if (wclaimed) { const uint32_t tot = atomicAdd(&s_ctl[C_CLAIMED], wclaimed) + wclaimed; if (tot > a.limit) s_ctl[C_ABORT] = 1u; }
                        atomicAdd(&s_ctl[C_DONE], 1u);
                    }
                    __syncwarp();
                }
                if (have && a.use_tma) { mbar_wait(bar, parity); parity ^= 1u; }      // left on abort: a copy for a batch this warp no longer takes
                __syncwarp();
            }
            __syncthreads();

            if (vs_ctl[C_ABORT]) {
                // the table filled up: forget this pass, split the sub-range by the next hash bits
                for (uint32_t i = tid; i < S; i += NT) { s_keys[i] = EMPTY_KEY; s_cnt[i] = 0; }
                if (tid == 0) {
                    // batches done when the limit was hit -> how many parts the sub-range needs (x1.5 margin)
                    const uint32_t done = max(1u, s_ctl[C_DONE]);
                    uint32_t F = 2;
                    while (F < BC_MAXP && (uint64_t)F * done * 2 < (uint64_t)n_batches_total * 3) F <<= 1;
                    while (P * F > BC_MAXP && F > 1) F >>= 1;
                    if (F < 2) {
                        const uint32_t i = atomicAdd(&a.ctl->n_heavy, 1u);
                        if (i < a.heavy_cap) { HeavyEnt e; e.bin = bin; e.p = (uint16_t)p; e.P = (uint16_t)P; a.heavy[i] = e; } else a.ctl->heavy_overflow = 1u;
                        atomicAdd(&a.ctl->heavy_recs, (unsigned long long)n_recs);
                    } else {
                        uint32_t spn = s_ctl[C_SP];
                        for (uint32_t i = 0; i < F; i++) s_stack[spn++] = (p + i * P) | ((P * F) << 16);
                        s_ctl[C_SP] = spn;
                        atomicAdd(&a.ctl->n_split, 1u);
                    }
                }
                __syncthreads();
                continue;
            }

            // ---- sweep: histogram of all entries, compaction of count > thr, clear for the next pass.  A slot is occupied
            // iff its count is non-zero (whoever claims a slot increments it), so phase 1 only reads the counts.
            uint32_t good_mask = 0, n_occ = 0, h1 = 0, h2 = 0;
#pragma unroll
            for (uint32_t u = 0; u < U; u++) {
                const uint32_t c = lds_u32(cnt_a + (u * NT + tid) * 4u);
                const uint32_t cc = c < MAX_COUNT ? c : MAX_COUNT;
                n_occ += c != 0u ? 1u : 0u;
                h1 += cc == 1u ? 1u : 0u;
                h2 += cc == 2u ? 1u : 0u;
                if (cc > 2u) { if (cc < (uint32_t)BC_HIST) atomicAdd(&s_hist[cc], 1u); else atomicAdd(&a.hist[cc], 1ULL); }
                if (cc > a.thr) good_mask |= 1u << u;
            }
            h1 = __reduce_add_sync(FULL, h1); h2 = __reduce_add_sync(FULL, h2);
            if (lane == 0) { if (h1) atomicAdd(&s_hist[1], h1); if (h2) atomicAdd(&s_hist[2], h2); }
            occ_total += n_occ;
            const uint32_t mine = __popc(good_mask);
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= (uint32_t)o) incl += v; }
            if (lane == 31) s_wgood[warp] = incl;
            __syncthreads();
            uint32_t before = 0, total_good = 0;
#pragma unroll
            for (uint32_t wv = 0; wv < NW; wv++) { const uint32_t c = s_wgood[wv]; if (wv < warp) before += c; total_good += c; }
            if (tid == 0 && total_good) *s_base = atomicAdd(&a.ctr->n_good, (unsigned long long)total_good);
            __syncthreads();
            uint64_t at = total_good ? *s_base + before + incl - mine : 0;
#pragma unroll
            for (uint32_t u = 0; u < U; u++) {
                const uint32_t slot = u * NT + tid;
                if ((good_mask >> u) & 1u) {
                    if (at < a.out_cap) {
                        const uint32_t c = lds_u32(cnt_a + slot * 4u);
                        a.out_keys[at] = lds_u64(keys_a + slot * 8u);
                        a.out_counts[at] = (uint16_t)(c < MAX_COUNT ? c : MAX_COUNT);
                    }
                    at++;
                }
                sts_u64(keys_a + slot * 8u, EMPTY_KEY); sts_u32(cnt_a + slot * 4u, 0u);
            }
            __syncthreads();
        }
        __syncthreads();
    }
    __syncthreads();
    for (uint32_t i = tid; i < BC_HIST; i += NT) if (s_hist[i]) atomicAdd(&a.hist[i], (unsigned long long)s_hist[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) occ_total += __shfl_xor_sync(FULL, occ_total, o);
    if (lane == 0 && occ_total) atomicAdd(&a.ctr->bc_distinct, occ_total);
    if (tid == 0 && recs_total) atomicAdd(&a.ctl->total_recs, recs_total);
}

// ------------------------------------------------------------------------------------------
// Heavy (bin, sub-range) entries and the overflow list: counted into the global table.  Placement is a
// function of the key alone (TableGeom: plain hash for the residual table of a bin-local count, the
// minimizer regions when a sample falls back to the region-blocked table for good).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
drain_heavy_kernel(BinSrc src, const HeavyEnt *__restrict__ heavy, uint32_t n_heavy, uint32_t blocks_per_ent, int k,
                   Slot *__restrict__ tab, TableGeom g, Counters *__restrict__ ctr) {
    __shared__ uint4 s_rec[8][32];
    __shared__ uint32_t s_pre[8][33];
    const uint32_t e = blockIdx.x / blocks_per_ent, sub = blockIdx.x % blocks_per_ent;
    if (e >= n_heavy) return;
    const HeavyEnt ent = heavy[e];
    const uint32_t p = ent.p, P = ent.P;
    uint32_t claimed = 0;
    auto put = [&](uint64_t key, uint32_t) {
        if (P > 1 && (bc_hash(key) & (P - 1u)) != p) return;
        claimed += placed_upsert_at<true>(tab, g.cap, g.region_shift, g.minimizer ? g.win : 0u, geom_home(g, key), key, 1u) ? 1u : 0u;
    };
    const uint32_t warp = threadIdx.x >> 5;
    for (uint32_t sj = 0; sj < src.n_src; sj++) {
        const uint32_t s = (sj + src.rot) % src.n_src;
        const uint64_t n = src.cursor[s][src.seg0 + ent.bin];
        const uint64_t n_seg = n > src.seg_cap ? src.seg_cap : n;
        skm_expand_records(src.recs[s] + (uint64_t)(src.seg0 + ent.bin) * src.seg_cap, n_seg, (uint64_t)sub * 256, (uint64_t)blocks_per_ent * 256, k, s_rec, s_pre, put);
        // overflow chunks: chunk j goes to warp j of the entry's CTAs (skm_expand_records hands warp w the records
        // [32 w, 32 w + 32) of its array, hence the shifted base pointer)
        const uint32_t n_over = (uint32_t)(n - n_seg), n_chunks = (n_over + 31) >> 5;
        for (uint32_t j = sub * 8 + warp; j < n_chunks; j += blocks_per_ent * 8) {
            uint32_t ch = 0;
            if ((threadIdx.x & 31) == 0) ch = ovf_chunk_of<false>(src.ovf_tags[s], src.ovf_chunks, src.seg0 + ent.bin, j, nullptr);
            ch = __shfl_sync(0xffffffffu, ch, 0);
            if (ch == OVF_NO_CHUNK) continue;
            const uint32_t cnt = min(32u, n_over - j * 32);
            skm_expand_records(src.ovf[s] + (size_t)ch * 32 - (size_t)warp * 32, (uint64_t)warp * 32 + cnt, 0, ~0ull >> 2, k, s_rec, s_pre, put);
        }
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if ((threadIdx.x & 31) == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

// every bin of the shard as a heavy entry (a sample that leaves the bin-local mode drains all its bins into the table)
__global__ void __launch_bounds__(256)
heavy_all_bins_kernel(HeavyEnt *__restrict__ heavy, uint32_t n_bins) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_bins; i += gridDim.x * blockDim.x) { HeavyEnt e; e.bin = i; e.p = 0; e.P = 1; heavy[i] = e; }
}

}  // namespace mfkc
