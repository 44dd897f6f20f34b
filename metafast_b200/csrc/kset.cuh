// kset.cuh -- device kernels of the (k-mer -> short) map algebra behind MetaFast's .kmers.bin consumers
// (SURVEY.md 8f rank 1): kmers-filter, unique-kmers-multi, kmers-samples-counter.
//
// The reference keeps BigLong2ShortHashMaps and walks them entry by entry
// (src/tools/KmersFilter.java:94-110, src/tools/UniqueKmersMultipleSamplesFinder.java:97-158,
// src/tools/KmersSamplesCounter.java:90-119, src/io/IOUtils.java:101-123,237-258,369-401).  Here a map is a
// key-sorted array pair on the device (keys u64, values = the 16-bit pattern of the Java short in a u32), so that
//   loadKmers          = parse + radix sort + weighted run-length (addAndBound of positive values = clamped sum)
//   put(get + x)       = concatenate (dst, src) + stable sort + pair combine          (kset_combine_kernel)
//   get / getWithZero  = binary search                                                  (kset_lower_bound)
//   filterAndPrintKmers = flag + compaction, records already in ascending key order.
#pragma once
#include "kernels.cuh"

namespace mfkc {

enum { KSET_ADD = 0, KSET_INC = 1, KSET_ZERO = 2 };

__device__ __forceinline__ int kset_short(uint32_t v) { return (int)(short)(v & 0xFFFFu); }
// Long2ShortHashMap.getWithZero ([itmo]/structures/map/Long2ShortHashMap.java:178-183): a stored -1 reads as 0 too
__device__ __forceinline__ int kset_gz(uint32_t v) { const int s = kset_short(v); return s == -1 ? 0 : s; }

// Kmers2HMWorker.processKmer (src/io/IOUtils.java:249-257): keep records with freq > threshold
__global__ void __launch_bounds__(256)
kset_parse_kernel(const uint8_t *__restrict__ recs, uint64_t n, int threshold, unsigned long long *__restrict__ keys,
                  uint32_t *__restrict__ vals, unsigned long long *__restrict__ cursor) {
    for (uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x; t0 < n; t0 += (uint64_t)gridDim.x * blockDim.x) {   // warp-uniform trip count
        const uint64_t t = t0 + threadIdx.x;
        unsigned long long key = 0; int freq = 0;
        bool keep = false;
        if (t < n) { load_record(recs + 10 * t, key, freq); keep = freq > threshold; }
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (!m) continue;
        const int leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if ((int)lane_id() == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (keep) {
            const uint64_t at = base + __popc(m & lanemask_lt());
            keys[at] = key;
            vals[at] = freq > 0 ? (uint32_t)freq : 0u;
        }
    }
}

__device__ __forceinline__ uint64_t kset_lower_bound(const unsigned long long *__restrict__ keys, uint64_t n, unsigned long long key) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// tag the entries of one side for the union sort: payload = value | side << 16; `only_above` drops nothing here, the
// caller compacts src first (kset_flag_kernel) when a threshold applies
__global__ void __launch_bounds__(256)
kset_tag_kernel(uint32_t *__restrict__ vals, uint64_t n, uint32_t side) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        vals[i] = (vals[i] & 0xFFFFu) | (side << 16);
}

// After the stable sort of (dst entries, then src entries): every key appears once or twice (dst first).  Writes one
// output entry per run head (positions from the rle_mark scan):
//   KSET_ADD: value = (short)(getWithZero(dst) + src)     UniqueKmersMultipleSamplesFinder.java:107-108
//   KSET_INC: value = (short)(getWithZero(dst) + 1)       UniqueKmersMultipleSamplesFinder.java:109, KmersSamplesCounter.java:103-105
__global__ void __launch_bounds__(256)
kset_combine_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, int op,
                    const unsigned long long *__restrict__ run_base, unsigned long long *__restrict__ out_keys,
                    uint32_t *__restrict__ out_vals) {
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_base;
    const uint64_t per_block = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = per_block * blockIdx.x;
    const uint64_t hi = lo + per_block < n ? lo + per_block : n;
    if (threadIdx.x == 0) s_base = run_base[blockIdx.x];
    __syncthreads();
    for (uint64_t start = lo; start < hi; start += blockDim.x) {
        const uint64_t i = start + threadIdx.x;
        const bool head = i < hi && (i == 0 || keys[i] != keys[i - 1]);
        const uint32_t m = __ballot_sync(0xffffffffu, head);
        if (lane_id() == 0) s_warp[threadIdx.x >> 5] = __popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int wv = 0; wv < 8; wv++) { const uint32_t c = s_warp[wv]; if (wv < (int)(threadIdx.x >> 5)) before += c; total += c; }
        if (head) {
            const uint64_t at = s_base + before + __popc(m & lanemask_lt());
            const unsigned long long key = keys[i];
            const uint32_t v0 = vals[i];
            const bool first_is_dst = (v0 >> 16) == 0;
            const bool pair = i + 1 < n && keys[i + 1] == key;
            uint32_t out;
            if (first_is_dst && !pair) out = v0 & 0xFFFFu;                              // untouched dst entry
            else {
                const int have = first_is_dst ? kset_gz(v0) : 0;                        // absent: get() = -1 -> 0
                const int add = op == KSET_INC ? 1 : kset_short(first_is_dst ? vals[i + 1] : v0);
                out = (uint32_t)(have + add) & 0xFFFFu;                                 // (short)(a + b): wraps
            }
            out_keys[at] = key;
            out_vals[at] = out;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
}

// KSET_ZERO, in place: if (src[key] > thr && dst.get(key) > thr) dst[key] = 0   (UniqueKmersMultipleSamplesFinder.java:127-129)
__global__ void __launch_bounds__(256)
kset_zero_kernel(const unsigned long long *__restrict__ dkeys, uint32_t *__restrict__ dvals, uint64_t dn,
                 const unsigned long long *__restrict__ skeys, const uint32_t *__restrict__ svals, uint64_t sn, int thr) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dn; i += (uint64_t)gridDim.x * blockDim.x) {
        if (kset_short(dvals[i]) <= thr) continue;
        const uint64_t p = kset_lower_bound(skeys, sn, dkeys[i]);
        if (p < sn && skeys[p] == dkeys[i] && kset_short(svals[p]) > thr) dvals[i] = 0u;
    }
}

// flags[i] = value > threshold && (no filter || filter.getWithZero(key) > filter_threshold)   (src/io/IOUtils.java:113)
__global__ void __launch_bounds__(256)
kset_flag_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, int threshold,
                 const unsigned long long *__restrict__ fkeys, const uint32_t *__restrict__ fvals, uint64_t fn, int has_filter,
                 int filter_threshold, uint8_t *__restrict__ flags) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        bool ok = kset_short(vals[i]) > threshold;
        if (ok && has_filter) {
            const uint64_t p = kset_lower_bound(fkeys, fn, keys[i]);
            const int fv = (p < fn && fkeys[p] == keys[i]) ? kset_gz(fvals[p]) : 0;
            ok = fv > filter_threshold;
        }
        flags[i] = ok ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256)
kset_flag_count_kernel(const uint8_t *__restrict__ flags, uint64_t n, unsigned long long *__restrict__ blk) {
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint64_t per_block = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = per_block * blockIdx.x;
    const uint64_t hi = lo + per_block < n ? lo + per_block : n;
    uint32_t local = 0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) local += flags[i];
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane_id() == 0 && local) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) blk[blockIdx.x] = s_cnt;
}

// order-preserving compaction of the flagged entries into (key, value | tag) arrays (V = u16 for records, u32 for the union sort)
template <typename V>
__global__ void __launch_bounds__(256)
kset_flag_write_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, const uint8_t *__restrict__ flags,
                       uint64_t n, const unsigned long long *__restrict__ blk_base, unsigned long long *__restrict__ out_keys,
                       V *__restrict__ out_vals, uint32_t tag) {
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_base;
    const uint64_t per_block = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = per_block * blockIdx.x;
    const uint64_t hi = lo + per_block < n ? lo + per_block : n;
    if (threadIdx.x == 0) s_base = blk_base[blockIdx.x];
    __syncthreads();
    for (uint64_t start = lo; start < hi; start += blockDim.x) {
        const uint64_t i = start + threadIdx.x;
        const bool good = i < hi && flags[i];
        const uint32_t m = __ballot_sync(0xffffffffu, good);
        if (lane_id() == 0) s_warp[threadIdx.x >> 5] = __popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int wv = 0; wv < 8; wv++) { const uint32_t c = s_warp[wv]; if (wv < (int)(threadIdx.x >> 5)) before += c; total += c; }
        if (good) {
            const uint64_t at = s_base + before + __popc(m & lanemask_lt());
            out_keys[at] = keys[i];
            out_vals[at] = (V)((vals[i] & 0xFFFFu) | tag);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
}

// hist[value]++ over all entries (QuickQuantitativeStatistics; values outside 0..32767 cannot be indexed in the reference either)
__global__ void __launch_bounds__(256)
kset_hist_kernel(const uint32_t *__restrict__ vals, uint64_t n, unsigned long long *__restrict__ hist) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int v = kset_short(vals[i]);
        if (v >= 0) atomicAdd(&hist[v], 1ULL);
    }
}

__global__ void __launch_bounds__(256)
kset_fill_kernel(uint32_t *__restrict__ vals, uint64_t n, uint32_t v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) vals[i] = v;
}

}  // namespace mfkc

// ------------------------------------------------------------------------------------------
// seq-builder: simple paths ("sequences") of the de Bruijn graph of the k-mers with count >
// threshold -- SequencesFinders.thresholdStrategy / AddSequencesShiftingRightTask (src/algo/SequencesFinders.java:13-31,
// src/algo/AddSequencesShiftingRightTask.java:39-123) and HashMapOperations.get{Left,Right}Nucleotide
// (src/algo/HashMapOperations.java:13-47).  The map is the key-sorted array pair of a mfkc_kset plus a hash index over
// it (key -> value) for the neighbour queries; one thread owns one map entry and tries both orientations, forward
// first, exactly like one iteration of the reference's entry loop (which is what makes its `used` rule deterministic).
// ------------------------------------------------------------------------------------------
namespace mfkc {

struct SeqIndex { const Slot *tab; uint64_t cap; int k; int thr; };

__global__ void __launch_bounds__(256)
seq_index_build_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
                       Slot *__restrict__ tab, uint64_t cap) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[t];
        uint64_t i = home_slot(key, cap);
        for (;;) {                                            // keys are unique: claim the first free slot
            if (atomicCAS(&tab[i].key, EMPTY_KEY, key) == EMPTY_KEY) { tab[i].count = vals[t] & 0xFFFFu; break; }
            if (++i == cap) i = 0;
        }
    }
}

// hm.get(key): the stored short, -1 when absent ([itmo]/structures/map/Long2ShortHashMap.java:160-175)
__device__ __forceinline__ int seq_get(const SeqIndex &ix, unsigned long long key) {
    uint64_t i = home_slot(key, ix.cap);
    for (;;) {
        const ulonglong2 s = ld_cg_u64x2(&ix.tab[i]);
        if (s.x == key) return kset_short((uint32_t)s.y);
        if (s.x == EMPTY_KEY) return -1;
        if (++i == ix.cap) i = 0;
    }
}
__device__ __forceinline__ unsigned long long seq_canon(unsigned long long fw, int k) {
    const unsigned long long rc = revcomp64(fw, k);
    return fw < rc ? fw : rc;
}
__device__ __forceinline__ unsigned long long seq_shift_right(unsigned long long fw, uint32_t nuc, int k) {
    return ((fw << 2) | nuc) & (k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull));
}
__device__ __forceinline__ unsigned long long seq_shift_left(unsigned long long fw, uint32_t nuc, int k) {
    return (fw >> 2) | ((unsigned long long)nuc << (2 * k - 2));
}
// unique extension nucleotide, -1 = none, -2 = several
__device__ __forceinline__ int seq_left_nuc(const SeqIndex &ix, unsigned long long fw) {
    int ans = -1;
#pragma unroll
    for (uint32_t nuc = 0; nuc < 4; nuc++)
        if (seq_get(ix, seq_canon(seq_shift_left(fw, nuc, ix.k), ix.k)) > ix.thr) { if (ans > -1) return -2; ans = (int)nuc; }
    return ans;
}
__device__ __forceinline__ int seq_right_nuc(const SeqIndex &ix, unsigned long long fw) {
    int ans = -1;
#pragma unroll
    for (uint32_t nuc = 0; nuc < 4; nuc++)
        if (seq_get(ix, seq_canon(seq_shift_right(fw, nuc, ix.k), ix.k)) > ix.thr) { if (ans > -1) return -2; ans = (int)nuc; }
    return ans;
}

struct SeqRecord {                 // one accepted sequence
    unsigned long long start_fw;   // its first k-mer, as read (directed)
    unsigned long long order;      // entry index * 2 + orientation: the deterministic output order
    uint32_t length, av_weight, min_weight, max_weight;
};

// processSequence (AddSequencesShiftingRightTask.java:75-122) without materialising the bases; out_bases != nullptr
// writes them ('A','G','C','T' for codes 0..3, ShortKmer.toString)
__device__ __forceinline__ void seq_walk(const SeqIndex &ix, unsigned long long fw, uint32_t &length, unsigned long long &weight,
                                         int &lo, int &hi, unsigned long long &end_fw, char *out_bases) {
    const int k = ix.k;
    int value = seq_get(ix, seq_canon(fw, k));
    if (value == -1) value = 0;                              // getWithZero
    weight = (unsigned long long)(long long)value; lo = hi = value;
    length = (uint32_t)k;
    if (out_bases)
        for (int i = 0; i < k; i++) out_bases[i] = "AGCT"[(fw >> (2 * (k - 1 - i))) & 3ull];
    unsigned long long cur = fw;
    for (;;) {
        const int r = seq_right_nuc(ix, cur);
        if (r < 0) break;
        const unsigned long long nxt = seq_shift_right(cur, (uint32_t)r, k);
        if (seq_left_nuc(ix, nxt) < 0) break;
        cur = nxt;
        if (out_bases) out_bases[length] = "AGCT"[r];
        length++;
        value = seq_get(ix, seq_canon(cur, k));
        if (value == -1) value = 0;
        weight += (unsigned long long)(long long)value;
        lo = value < lo ? value : lo; hi = value > hi ? value : hi;
    }
    end_fw = cur;
}

// One thread per map entry: both orientations (forward first).  recs == nullptr: count only.
__global__ void __launch_bounds__(256)
seq_find_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, SeqIndex ix,
                uint32_t len_threshold, SeqRecord *__restrict__ recs, unsigned long long *__restrict__ cursor /* [0] sequences, [1] bases */) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        if (kset_short(vals[t]) <= ix.thr) continue;
        const unsigned long long key = keys[t];
        bool used = false;                                    // `used` set of the reference, restricted to this key
#pragma unroll 1
        for (int o = 0; o < 2; o++) {
            const unsigned long long fw = o ? revcomp64(key, ix.k) : key;
            const int nuc = seq_left_nuc(ix, fw);
            bool is_left = nuc < 0;
            if (!is_left) is_left = seq_right_nuc(ix, seq_shift_left(fw, (uint32_t)nuc, ix.k)) < 0;
            if (!is_left) continue;
            uint32_t length; unsigned long long weight, end_fw; int lo, hi;
            seq_walk(ix, fw, length, weight, lo, hi, end_fw, nullptr);
            if (length < len_threshold) continue;
            const unsigned long long st = seq_canon(fw, ix.k), en = seq_canon(end_fw, ix.k);
            if (st > en) continue;                            // the walk from the other end prints it
            if (st == en) { if (used) continue; used = true; }
            const unsigned long long at = atomicAdd(&cursor[0], 1ULL);
            atomicAdd(&cursor[1], (unsigned long long)length);
            if (recs) {
                SeqRecord r;
                r.start_fw = fw; r.order = 2 * t + (unsigned long long)o; r.length = length;
                r.av_weight = (uint32_t)(weight / (unsigned long long)(length - ix.k + 1));
                r.min_weight = (uint32_t)lo; r.max_weight = (uint32_t)hi;
                recs[at] = r;
            }
        }
    }
}

// One thread per accepted sequence (records sorted by `order`, base offsets prefix-summed on the host): walk again, write bases
__global__ void __launch_bounds__(128)
seq_write_kernel(const SeqRecord *__restrict__ recs, const unsigned long long *__restrict__ base_off, uint64_t n_seq, SeqIndex ix,
                 char *__restrict__ bases) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_seq; t += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t length; unsigned long long weight, end_fw; int lo, hi;
        seq_walk(ix, recs[t].start_fw, length, weight, lo, hi, end_fw, bases + base_off[t]);
    }
}

}  // namespace mfkc
