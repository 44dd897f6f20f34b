// kernels128.cuh -- the counting path for long k-mers, 32 <= k <= 63 (128-bit keys).
//
// The reference stops at k = 31 (src/tools/KmersCounterMain.java:70-73); this is the natural
// extension SURVEY.md 8c defines: key = min(fw, rc) as unsigned 2k-bit integers in the same
// A0 G1 C2 T3 encoding, first base most significant; record = 16-byte BE key + 2-byte BE count.
// Validated against the Python-int restatement kept with the tests only -- parity
// unpinned.  Same design as the 64-bit path: super-k-mer staging by minimizer region, L2-resident
// drain, 128-bit atomicCAS (ATOMG.E.CAS.128) to claim a slot.
#pragma once
#include "kernels.cuh"

namespace mfkc {

struct K128 { unsigned long long lo, hi; };             // value = hi * 2^64 + lo
__device__ __forceinline__ bool k128_less(const K128 &a, const K128 &b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
__device__ __forceinline__ bool k128_eq(const K128 &a, const K128 &b) { return a.hi == b.hi && a.lo == b.lo; }
__host__ __device__ __forceinline__ uint64_t mix128(unsigned long long lo, unsigned long long hi) { return mix64(lo ^ mix64(hi + 0x9E3779B97F4A7C15ULL)); }

// One slot = one 32-byte sector: key (16 B, the unit of the 128-bit CAS), count.
struct __align__(32) Slot128 {
    unsigned long long lo, hi;                          // empty = all ones (legal keys are < 2^126)
    uint32_t count;
    uint32_t pad[3];
};
static_assert(sizeof(Slot128) == 32, "Slot128 must be one sector");

__device__ __forceinline__ K128 cas128(void *addr, K128 cmp, K128 val) {
    K128 old;
    asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
                 "atom.global.cas.b128 o, [%6], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
                 : "=l"(old.lo), "=l"(old.hi) : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "l"(addr) : "memory");
    return old;
}

// one slot of a probe sequence: returns true when the key was found or placed there (count = sat_add(count, inc))
__device__ __forceinline__ bool slot128_try(Slot128 *__restrict__ tab, uint64_t i, K128 key, uint32_t inc, bool &claimed) {
    const K128 empty{~0ull, ~0ull};
    const ulonglong2 kk = ld_cg_u64x2(&tab[i]);                   // lo, hi
    K128 cur{kk.x, kk.y};
    bool mine = false;
    if (k128_eq(cur, empty)) {
        cur = cas128(&tab[i], empty, key);
        if (k128_eq(cur, empty)) { mine = true; cur = key; }
    }
    if (!k128_eq(cur, key)) return false;
    if (inc == 1) {
        const uint32_t c = *(volatile uint32_t *)&tab[i].count;
        if (c < MAX_COUNT) atomicAdd(&tab[i].count, 1u);
    } else {
        uint32_t old = *(volatile uint32_t *)&tab[i].count;
        for (;;) {
            if (old >= MAX_COUNT) break;
            uint32_t nv = old + inc; if (nv > MAX_COUNT || nv < old) nv = MAX_COUNT;
            const uint32_t seen = atomicCAS(&tab[i].count, old, nv);
            if (seen == old) break;
            old = seen;
        }
    }
    claimed = mine;
    return true;
}

// Windowed placement, as for the 64-bit tables (placed_upsert_at in kernels.cuh): primary window of SMEM_WIN slots from
// the home slot, wrapping inside the region; then windows at fresh uniform positions.  Returns true when a slot was claimed.
__device__ __forceinline__ bool table128_upsert_at(Slot128 *__restrict__ tab, uint64_t cap, int region_shift, uint64_t home, K128 key, uint32_t inc) {
    const uint64_t mask = (1ull << region_shift) - 1ull;
    const uint64_t rbase = home & ~mask;
    uint64_t off = home & mask;
    const uint32_t steps = (uint64_t)SMEM_WIN < mask + 1 ? SMEM_WIN : (uint32_t)(mask + 1);
    bool claimed = false;
    for (uint32_t st = 0; st < steps; st++, off = (off + 1) & mask)
        if (slot128_try(tab, rbase | off, key, inc, claimed)) return claimed;
    for (uint32_t a = 1;; a++) {
        uint64_t i = mulhi64(mix128(key.lo ^ (SECONDARY_SALT * a), key.hi), cap);
        for (uint32_t st = 0; st < SMEM_WIN; st++) {
            if (slot128_try(tab, i, key, inc, claimed)) return claimed;
            if (++i == cap) i = 0;
        }
    }
}

__global__ void __launch_bounds__(256)
table128_clear_kernel(Slot128 *__restrict__ tab, uint64_t cap) {
    const uint4 e0 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), e1 = make_uint4(0u, 0u, 0u, 0u);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * cap; i += (uint64_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint4 *>(tab)[i] = (i & 1) ? e1 : e0;
}

// ---- minimizer of a 128-bit key (slow path: rehash)
__device__ __forceinline__ uint32_t minhash_of_key128(K128 key, int k) {
    const int m = minimizer_len(k);
    const uint32_t mask = (1u << (2 * m)) - 1u;
    uint32_t best = 0xFFFFFFFFu, fw = 0, rc = 0;
    for (int i = 0; i < k; i++) {
        const int sh = 2 * (k - 1 - i);
        const uint32_t c = (uint32_t)(sh >= 64 ? key.hi >> (sh - 64) : key.lo >> sh) & 3u;
        fw = ((fw << 2) | c) & mask;
        rc = (rc >> 2) | ((3u - c) << (2 * m - 2));
        if (i >= m - 1) { const uint32_t h = mmer_hash(fw < rc ? fw : rc); best = h < best ? h : best; }
    }
    return best;
}

struct TableGeom128 { uint64_t cap; uint32_t n_regions; int region_shift; int k; };
__device__ __forceinline__ uint64_t home128(uint32_t region, int region_shift, K128 key) {
    return ((uint64_t)region << region_shift) | (mix128(key.lo, key.hi) & ((1ull << region_shift) - 1ull));
}

__global__ void __launch_bounds__(256)
rehash128_kernel(const Slot128 *__restrict__ old_tab, uint64_t old_cap, Slot128 *__restrict__ new_tab, TableGeom128 g) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < old_cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const ulonglong2 kk = ld_cg_u64x2(&old_tab[i]);
        if (kk.x == ~0ull && kk.y == ~0ull) continue;
        const K128 key{kk.x, kk.y};
        const uint32_t c = old_tab[i].count;
        const uint32_t region = region_of_minhash(minhash_of_key128(key, g.k), g.n_regions);
        table128_upsert_at(new_tab, g.cap, g.region_shift, home128(region, g.region_shift, key), key, c < MAX_COUNT ? c : MAX_COUNT);
    }
}

// ---- front end: 5-word window (80 bases) + 128 boundary flags
struct TileWord5 {
    uint32_t w[5];
    unsigned long long f_lo, f_hi;     // boundary flags of positions [16w, 16w+128)
    long long limit;
    bool active;
};
template <int NT>
__device__ __forceinline__ TileWord5 load_tile_word5(const uint8_t *__restrict__ bases, uint64_t n_bases,
                                                     const uint32_t *__restrict__ flags, uint64_t tile, int k,
                                                     uint32_t *s_words /* NT+4 */, uint32_t *s_flags /* NT/2+4 */, uint32_t &bad) {
    const uint32_t tid = threadIdx.x;
    const uint64_t n_flag_words = (n_bases + 31) >> 5;
    const uint64_t w_base = tile * NT;
#pragma unroll
    for (int rep = 0; rep < 2; rep++) {
        if (rep == 1 && tid >= 4) break;
        const uint32_t slot = rep ? NT + tid : tid;
        const uint64_t w = w_base + slot;
        uint32_t word = 0;
        const uint64_t b0 = w << 4;
        if (b0 + 16 <= n_bases) {
            const uint4 v = ld_nc_u128(bases + b0);
            word = pack16(v);
            bad |= bad4(v.x) | bad4(v.y) | bad4(v.z) | bad4(v.w);
        } else if (b0 < n_bases) {
            for (uint32_t j = 0; j < 16 && b0 + j < n_bases; j++) {
                const uint32_t c = bases[b0 + j];
                bad |= bad4(c | 0x41414100u);
                word |= pack4(c) >> 6 << (30 - 2 * j);
            }
        }
        s_words[slot] = word;
    }
    if (tid < NT / 2 + 4) {
        const uint64_t fw = (w_base >> 1) + tid;
        s_flags[tid] = fw < n_flag_words ? flags[fw] : 0u;
    }
    __syncthreads();
    TileWord5 t;
    const uint64_t w = w_base + tid;
    t.active = (w << 4) < n_bases;
#pragma unroll
    for (int i = 0; i < 5; i++) t.w[i] = s_words[tid + i];
    const uint32_t b = tid >> 1;
    const uint64_t a0 = (uint64_t)s_flags[b] | ((uint64_t)s_flags[b + 1] << 32);
    const uint64_t a1 = (uint64_t)s_flags[b + 2] | ((uint64_t)s_flags[b + 3] << 32);
    const uint64_t a2 = (uint64_t)s_flags[b + 4];                                  // only its low 16 bits can matter (odd threads)
    if (tid & 1) { t.f_lo = (a0 >> 16) | (a1 << 48); t.f_hi = (a1 >> 16) | (a2 << 48); }
    else { t.f_lo = a0; t.f_hi = a1; }
    t.limit = (long long)n_bases - k - (long long)(w << 4);
    return t;
}

// validity of the 16 start offsets: no boundary flag in [p, p+k-2], p + k <= n_bases
__device__ __forceinline__ uint32_t valid_starts128(unsigned long long f_lo, unsigned long long f_hi, long long limit, int k) {
    uint32_t valid = 0;
    const int span = k - 1;                                  // 31..62 flag bits
    const unsigned long long mask = span >= 64 ? ~0ull : ((1ull << span) - 1ull);
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const unsigned long long win = j ? ((f_lo >> j) | (f_hi << (64 - j))) : f_lo;
        valid |= (((win & mask) == 0 && (long long)j <= limit) ? 1u : 0u) << j;
    }
    return valid;
}

__device__ __forceinline__ unsigned long long revpairs64(unsigned long long x) {
    x = __brevll(x);
    return ((x & 0x5555555555555555ULL) << 1) | ((x >> 1) & 0x5555555555555555ULL);
}
__device__ __forceinline__ K128 shr128(K128 a, int s) {      // 0 < s <= 64
    K128 r;
    if (s >= 64) { r.lo = a.hi; r.hi = 0; }
    else { r.lo = (a.lo >> s) | (a.hi << (64 - s)); r.hi = a.hi >> s; }
    return r;
}
// canonical keys of the (up to 16) k-mers starting at base offsets 0..15 of a 5-word window
__device__ __forceinline__ void kmers128_of_word(const uint32_t (&w)[5], int k, K128 (&keys)[16]) {
    const int s = 128 - 2 * k;                               // 2..64
    const int top = 2 * k - 2;                               // 62..124
    K128 rc{0, 0};
#pragma unroll
    for (int j = 0; j < 16; j++) {
        uint32_t v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = j ? __funnelshift_l(w[i + 1], w[i], 2 * j) : w[i];
        const K128 x{((unsigned long long)v[2] << 32) | v[3], ((unsigned long long)v[0] << 32) | v[1]};
        const K128 fw = shr128(x, s);
        if (j == 0) {
            K128 r{~revpairs64(fw.hi), ~revpairs64(fw.lo)};  // reversed pairs: old lo becomes the new hi
            rc = shr128(r, s);
        } else {
            rc.lo = (rc.lo >> 2) | (rc.hi << 62);
            rc.hi >>= 2;
            const unsigned long long c = (~fw.lo) & 3ull;
            if (top >= 64) rc.hi |= c << (top - 64); else rc.lo |= c << top;
        }
        keys[j] = k128_less(fw, rc) ? fw : rc;
    }
}

// minimizer hash of the 16 k-mers starting in word 0 (w = k - m + 1 >= 21 m-mers per k-mer):
// window_j = min(suffix[j..15], middle[16..w-1], prefix[w..w-1+j])
__device__ __forceinline__ void minhash128_of_word(const uint32_t (&w5)[5], int k, uint32_t (&mh)[16]) {
    const int m = minimizer_len(k);
    const int wn = k - m + 1;
    const int rs = 32 - 2 * m;
    uint32_t W[6];
#pragma unroll
    for (int i = 0; i < 5; i++) W[i] = w5[i];
    W[5] = 0;
    auto h_at = [&](int q) {
        const int i = q >> 4, sh = 2 * (q & 15);
        const uint32_t v = sh ? __funnelshift_l(W[i + 1], W[i], sh) : W[i];
        const uint32_t fw = v >> rs;
        uint32_t x = __brev(fw);
        x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
        const uint32_t rc = (~x) >> rs;
        return mmer_hash(fw < rc ? fw : rc);
    };
    uint32_t suf[16];
    uint32_t run = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 15; j >= 0; j--) { run = min(run, h_at(j)); suf[j] = run; }
    uint32_t mid = 0xFFFFFFFFu;
    for (int q = 16; q < wn; q++) mid = min(mid, h_at(q));
    run = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (j) run = min(run, h_at(wn - 1 + j));
        mh[j] = min(min(suf[j], mid), run);
    }
}

// staging of 32-byte records: 5 words of bases (len - 1 in the low 4 bits of the 5th), minimizer hash
struct SkmStage128 {
    uint4 *recs;                  // 2 x uint4 per record; region r owns records [r*seg_cap, (r+1)*seg_cap)
    unsigned int *cursor;
    uint64_t seg_cap;
    uint32_t n_regions;
    int region_shift;
};

__device__ __noinline__ uint32_t skm128_count_direct(const uint32_t (&w)[5], uint32_t len, uint32_t region, int region_shift,
                                                     int k, Slot128 *__restrict__ tab, uint64_t cap) {
    K128 keys[16];
    kmers128_of_word(w, k, keys);
    uint32_t claimed = 0;
#pragma unroll
    for (uint32_t t = 0; t < 16; t++)
        if (t < len) claimed += table128_upsert_at(tab, cap, region_shift, home128(region, region_shift, keys[t]), keys[t], 1u) ? 1u : 0u;
    return claimed;
}

// MODE 0: bucket = table region of this GPU; MODE 2: peer-memory exchange, bucket = owner shard << log2B | coarse bucket
// (st.n_regions = shards, st.region_shift = log2B; runs are cut by minimizer hash; kmer_count[owner] += run length)
template <int MODE>
__global__ void __launch_bounds__(EX_THREADS)
extract_skm128_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, const uint32_t *__restrict__ flags,
                      int k, SkmStage128 st, Slot128 *__restrict__ tab, uint64_t cap, Counters *__restrict__ ctr,
                      unsigned long long *__restrict__ kmer_count) {
    __shared__ uint32_t s_words[EX_THREADS + 4];
    __shared__ uint32_t s_flags[EX_THREADS / 2 + 4];
    const uint64_t n_tiles = (((n_bases + 15) >> 4) + EX_THREADS - 1) / EX_THREADS;
    uint32_t claimed = 0, bad = 0, dropped = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileWord5 t = load_tile_word5<EX_THREADS>(bases, n_bases, flags, tile, k, s_words, s_flags, bad);
        const uint32_t valid = t.active ? valid_starts128(t.f_lo, t.f_hi, t.limit, k) : 0u;
        unsigned long long km_lo = 0, km_hi = 0;                // MODE 2: k-mers per owner of this tile, 16-bit fields
        if (valid) {
            uint32_t mh[16];
            minhash128_of_word(t.w, k, mh);
            uint32_t run_start = 0, run_key = 0, run_mh = 0;
            bool in_run = false;
#pragma unroll
            for (int j = 0; j <= 16; j++) {
                const bool v = j < 16 && ((valid >> j) & 1);
                const uint32_t mhj = mh[j < 16 ? j : 15];
                const uint32_t key = MODE == 2 ? mhj : (v ? region_of_minhash(mhj, st.n_regions) : 0xFFFFFFFFu);
                if (in_run && (!v || key != run_key)) {
                    const uint32_t len = (uint32_t)j - run_start;
                    const int sh = 2 * (int)run_start;
                    uint32_t r[5];
#pragma unroll
                    for (int i = 0; i < 5; i++) r[i] = sh ? __funnelshift_l(i < 4 ? t.w[i + 1] : 0u, t.w[i], sh) : t.w[i];
                    const uint32_t owner = MODE == 2 ? owner_of_minhash(run_mh, st.n_regions) : 0u;
                    const uint32_t bucket = MODE == 2 ? ((owner << st.region_shift) | coarse_of_minhash(run_mh, st.region_shift)) : run_key;
                    const uint32_t pos = atomicAdd(&st.cursor[bucket], 1u);
                    if (pos < st.seg_cap) {
                        uint4 *dst = st.recs + 2 * ((uint64_t)bucket * st.seg_cap + pos);
                        dst[0] = make_uint4(r[0], r[1], r[2], r[3]);
                        dst[1] = make_uint4((r[4] & ~15u) | (len - 1), run_mh, 0u, 0u);
                        if (MODE == 2) {
                            if (st.n_regions <= 8) { if (owner < 4) km_lo += (unsigned long long)len << (16 * owner); else km_hi += (unsigned long long)len << (16 * (owner - 4)); }
                            else atomicAdd(&kmer_count[owner], (unsigned long long)len);
                        }
                    } else if (MODE == 2) {
                        dropped++;                           // reported as an error by mfkc_flush
                    } else {
                        r[4] &= ~15u;
                        claimed += skm128_count_direct(r, len, run_key, st.region_shift, k, tab, cap);
                    }
                    in_run = false;
                }
                if (v && !in_run) { in_run = true; run_start = (uint32_t)j; run_key = key; run_mh = mhj; }
            }
        }
        if (MODE == 2 && st.n_regions <= 8) {                   // one atomic per (warp, owner) and tile: G hot addresses would serialise in L2
#pragma unroll
            for (int o = 16; o; o >>= 1) { km_lo += __shfl_xor_sync(0xffffffffu, km_lo, o); km_hi += __shfl_xor_sync(0xffffffffu, km_hi, o); }
            const uint32_t l = lane_id();
            if (l < st.n_regions) {
                const uint32_t c = (uint32_t)((l < 4 ? km_lo >> (16 * l) : km_hi >> (16 * (l - 4))) & 0xFFFFu);
                if (c) atomicAdd(&kmer_count[l], (unsigned long long)c);
            }
        }
        __syncthreads();
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
    if (__any_sync(0xffffffffu, bad != 0) && lane_id() == 0) atomicAdd(&ctr->bad_chars, 1ULL);
    if (MODE == 2 && dropped) atomicAdd(&ctr->overflow, (unsigned long long)dropped);
}

// canonical key of the k-mer that starts `off` bases into a 5-word window (the j-th key of kmers128_of_word, computed alone)
__device__ __forceinline__ K128 kmer128_at(const uint32_t (&w)[5], uint32_t off, int k) {
    const int s = 128 - 2 * k;
    uint32_t v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = __funnelshift_l(w[i + 1], w[i], 2 * off);          // off = 0: shift 0 returns w[i]
    const K128 x{((unsigned long long)v[2] << 32) | v[3], ((unsigned long long)v[0] << 32) | v[1]};
    const K128 fw = shr128(x, s);
    const K128 r{~revpairs64(fw.hi), ~revpairs64(fw.lo)};
    const K128 rc = shr128(r, s);
    return k128_less(fw, rc) ? fw : rc;
}

// Work-balanced expansion of 32-byte records (the 128-bit twin of skm_expand_records): a warp takes 32 records, prefix-sums
// their lengths and deals the k-mer instances out evenly; the next 32 records are in flight meanwhile.
template <class Put>
__device__ __forceinline__ void skm128_expand_records(const uint4 *__restrict__ recs, uint64_t n, uint64_t first, uint64_t stride, int k,
                                                      uint4 (*s_a)[32], uint4 (*s_b)[32], uint32_t (*s_pre)[33], Put put) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t base = first + (uint64_t)warp * 32;
    uint4 na = make_uint4(0u, 0u, 0u, 0u), nb = na;
    if (base + lane < n) { na = ld_nc_u128(&recs[2 * (base + lane)]); nb = ld_nc_u128(&recs[2 * (base + lane) + 1]); }
    for (; base < n; base += stride) {
        const uint64_t i = base + lane;
        const uint4 a = na, b = nb;
        const uint32_t len = i < n ? (b.x & 15u) + 1u : 0u;
        if (i + stride < n) { na = ld_nc_u128(&recs[2 * (i + stride)]); nb = ld_nc_u128(&recs[2 * (i + stride) + 1]); }
        uint32_t incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
        s_a[warp][lane] = a; s_b[warp][lane] = b;
        s_pre[warp][lane + 1] = incl;
        if (lane == 0) s_pre[warp][0] = 0;
        __syncwarp();
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        for (uint32_t t = lane; t < total; t += 32) {
            uint32_t lo = 0, hi = 31;
#pragma unroll
            for (int step = 0; step < 5; step++) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (s_pre[warp][mid] <= t) lo = mid; else hi = mid - 1;
            }
            const uint32_t off = t - s_pre[warp][lo];
            const uint4 qa = s_a[warp][lo], qb = s_b[warp][lo];
            const uint32_t w[5] = {qa.x, qa.y, qa.z, qa.w, qb.x & ~15u};
            put(kmer128_at(w, off, k), qb.y);
        }
        __syncwarp();
    }
}

// peer-memory drain for 128-bit keys (see drain_p2p_kernel): one 32-byte record per thread, read from the peers' staging
struct P2PPeers;
__global__ void __launch_bounds__(256)
drain_p2p128_kernel(const uint4 *const *__restrict__ peer_recs, const unsigned int *const *__restrict__ peer_cursor, uint32_t n_peers,
                    uint32_t me, int log2_buckets, uint64_t seg_cap, uint32_t bucket0, uint32_t blocks_per_bucket, int k,
                    Slot128 *__restrict__ tab, uint64_t cap, uint32_t n_regions, int region_shift, Counters *__restrict__ ctr) {
    const uint32_t bucket = bucket0 + blockIdx.x / blocks_per_bucket;
    const uint32_t sub = blockIdx.x % blocks_per_bucket;
    const uint64_t seg = ((uint64_t)me << log2_buckets) | bucket;
    __shared__ uint4 s_a[8][32], s_b[8][32];
    __shared__ uint32_t s_pre[8][33];
    uint32_t claimed = 0;
    for (uint32_t j = 0; j < n_peers; j++) {
        const uint32_t s = (me + j) % n_peers;
        uint64_t n = peer_cursor[s][seg];
        if (n > seg_cap) n = seg_cap;
        skm128_expand_records(peer_recs[s] + 2 * seg * seg_cap, n, (uint64_t)sub * 256, (uint64_t)blocks_per_bucket * 256, k, s_a, s_b, s_pre,
                              [&](K128 key, uint32_t mh) {
                                  const uint32_t region = region_of_minhash(mh, n_regions);
                                  claimed += table128_upsert_at(tab, cap, region_shift, home128(region, region_shift, key), key, 1u) ? 1u : 0u;
                              });
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

__global__ void __launch_bounds__(256)
drain_skm128_kernel(SkmStage128 st, uint32_t blocks_per_region, int k, Slot128 *__restrict__ tab, uint64_t cap,
                    Counters *__restrict__ ctr) {
    const uint32_t region = blockIdx.x / blocks_per_region;
    const uint32_t sub = blockIdx.x % blocks_per_region;
    uint64_t n = st.cursor[region];
    if (n > st.seg_cap) n = st.seg_cap;
    const uint4 *__restrict__ recs = st.recs + 2 * (uint64_t)region * st.seg_cap;
    __shared__ uint4 s_a[8][32], s_b[8][32];
    __shared__ uint32_t s_pre[8][33];
    uint32_t claimed = 0;
    skm128_expand_records(recs, n, (uint64_t)sub * 256, (uint64_t)blocks_per_region * 256, k, s_a, s_b, s_pre, [&](K128 key, uint32_t) {
        claimed += table128_upsert_at(tab, cap, st.region_shift, home128(region, st.region_shift, key), key, 1u) ? 1u : 0u;
    });
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

// ---- emit: histogram + filter + compaction into (lo, {hi,count}) pairs
struct __align__(16) HiCount { unsigned long long other; uint32_t count; uint32_t pad; };

__global__ void __launch_bounds__(256)
table128_scan_kernel(const Slot128 *__restrict__ tab, uint64_t cap, uint32_t threshold, unsigned long long *__restrict__ hist,
                     unsigned long long *__restrict__ out_lo, HiCount *__restrict__ out_hc, uint64_t out_cap,
                     Counters *__restrict__ ctr) {
    __shared__ uint32_t s_hist[HIST_SMEM_BINS];
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_base;
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_iter = (cap + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t warp = threadIdx.x >> 5;
    for (uint64_t it = 0; it < n_iter; it++, i += stride) {
        bool good = false;
        unsigned long long lo = 0, hi = 0; uint32_t c = 0;
        if (i < cap) {
            const uint4 a = ld_nc_u128(&tab[i]);
            lo = ((unsigned long long)a.y << 32) | a.x; hi = ((unsigned long long)a.w << 32) | a.z;
            if (!(lo == ~0ull && hi == ~0ull)) {
                c = tab[i].count; if (c > MAX_COUNT) c = MAX_COUNT;
                if (c < HIST_SMEM_BINS) atomicAdd(&s_hist[c], 1u); else atomicAdd(&hist[c], 1ULL);
                good = c > threshold;
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, good);
        if (lane_id() == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int wv = 0; wv < 8; wv++) { const uint32_t cc = s_warp[wv]; if (wv < (int)warp) before += cc; total += cc; }
        if (threadIdx.x == 0 && total) s_base = atomicAdd(&ctr->n_good, (unsigned long long)total);
        __syncthreads();
        if (good) {
            const uint64_t at = s_base + before + __popc(m & lanemask_lt());
            if (at < out_cap) { out_lo[at] = lo; HiCount h; h.other = hi; h.count = c; h.pad = 0; out_hc[at] = h; }
        }
    }
    __syncthreads();
    for (int i2 = threadIdx.x; i2 < HIST_SMEM_BINS; i2 += blockDim.x)
        if (s_hist[i2]) atomicAdd(&hist[i2], (unsigned long long)s_hist[i2]);
}

// after the low word has been sorted: make the high word the sort key (payload keeps the low word)
__global__ void __launch_bounds__(256)
swap_key_words_kernel(unsigned long long *__restrict__ keys, HiCount *__restrict__ hc, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long k0 = keys[i];
        keys[i] = hc[i].other;
        hc[i].other = k0;
    }
}

// sorted (hi, {lo,count}) -> 18-byte big-endian records (16-byte key, 2-byte count)
__global__ void __launch_bounds__(256)
records128_kernel(const unsigned long long *__restrict__ hi, const HiCount *__restrict__ lc, uint64_t n,
                  uint16_t *__restrict__ out /* 9 x u16 per record */) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long h = hi[i], l = lc[i].other;
        const uint32_t c = lc[i].count;
        uint16_t *o = out + 9 * i;
        const uint32_t w[4] = {(uint32_t)(h >> 32), (uint32_t)h, (uint32_t)(l >> 32), (uint32_t)l};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            o[2 * q] = (uint16_t)__byte_perm(w[q], 0, 0x0023);
            o[2 * q + 1] = (uint16_t)__byte_perm(w[q], 0, 0x0001);
        }
        o[8] = (uint16_t)__byte_perm(c, 0, 0x0001);
    }
}

}  // namespace mfkc
