// kernels.cuh -- hand-written sm_100a kernels of the k-mer counting path (64-bit keys, k <= 31).
//
//   K0 mark_read_ends_kernel   read boundaries -> 1 bit per base, read/length statistics
//   K1+K2+K3 extract_kernel<Sink>  ASCII -> 2-bit pack (smem) -> canonical k-mers -> sink
//        sinks: table upsert (hash variant), append (sort variant), shard bucketing,
//               features-calculator presence accumulate
//   K3  rehash_kernel          table growth
//   K4  table_hist_kernel / table_compact_kernel / records_kernel   histogram, filter, BE records
//   K6  fc_* kernels           membership + accumulate, per-component reduce
//
// Reference semantics restated here:
//   [itmo]/dna/kmers/ShortKmer.java:54-56,68-71,122-149  (canonical rolling k-mers)
//   [itmo]/utils/KmerUtils.java:12-22                    (reverse complement)
//   [itmo]/structures/map/Long2ShortHashMap.java:119-157 (addAndBound = saturating count)
//   src/io/IOUtils.java:45-71, 577-588, 756-769, 806-825
#pragma once
#include "device_common.cuh"

namespace mfkc {

// ------------------------------------------------------------------------------------------
// device-side counters (one struct per context, lives in device memory)
// ------------------------------------------------------------------------------------------
struct Counters {
    unsigned long long distinct;      // successful slot claims (hm.size())
    unsigned long long kmers;         // k-mer instances extracted
    unsigned long long total_seq, good_seq, total_len, good_len;   // IOUtils.java:752-769
    unsigned long long bad_chars;     // bytes outside AaCcGgTt seen by the packer
    unsigned long long appended;      // sort variant / bucketing: keys written
    unsigned long long overflow;      // appended past capacity (keys dropped -> error)
    unsigned long long n_good;        // compaction cursor
    unsigned long long bc_distinct;   // bin-local counting: distinct k-mers counted in shared memory (bincount.cuh)
    unsigned long long pad[5];
};

// ------------------------------------------------------------------------------------------
// K0: read boundaries.  flags: 1 bit per base position (bit p&31 of word p>>5), set at the LAST
// base of every read, and at every base of reads that must yield nothing although they are
// >= k long (len < min_seq_len, IOUtils.java:761).  A k-mer starting at p is valid iff no flag
// is set in [p, p+k-2] and p+k <= n_bases; reads shorter than k are thereby skipped for free
// (ShortKmer.java:123,127).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mark_read_ends_kernel(const uint64_t *__restrict__ offsets, uint32_t n_reads, uint64_t n_bases,
                      int k, int min_len, int count_stats, uint32_t *__restrict__ flags,
                      Counters *__restrict__ ctr) {
    unsigned long long t_seq = 0, g_seq = 0, t_len = 0, g_len = 0, kmers = 0;
    const uint64_t base0 = offsets[0];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_reads; i += gridDim.x * blockDim.x) {
        uint64_t s = offsets[i] - base0, e = offsets[i + 1] - base0;
        if (e > n_bases) e = n_bases;
        if (s > e) s = e;
        const uint64_t len = e - s;
        t_seq++; t_len += len;
        const bool good = (long long)len >= (long long)min_len;
        if (good) { g_seq++; g_len += len; if (len >= (uint64_t)k) kmers += len - k + 1; }
        if (len == 0) continue;
        if (good || len < (uint64_t)k) {
            // k == 1: 1-mers never cross a boundary, and the flag window degenerates to the start
            // position itself (see kmers_of_word), so ordinary reads must stay unflagged
            if (k > 1) atomicOr(&flags[(e - 1) >> 5], 1u << ((e - 1) & 31));
        } else {
            for (uint64_t p = s; p < e; p++) atomicOr(&flags[p >> 5], 1u << (p & 31));
        }
    }
    if (!count_stats) return;
    // warp-reduce, one atomic per warp
    for (int o = 16; o; o >>= 1) {
        t_seq += __shfl_xor_sync(0xffffffffu, t_seq, o); g_seq += __shfl_xor_sync(0xffffffffu, g_seq, o);
        t_len += __shfl_xor_sync(0xffffffffu, t_len, o); g_len += __shfl_xor_sync(0xffffffffu, g_len, o);
        kmers += __shfl_xor_sync(0xffffffffu, kmers, o);
    }
    if (lane_id() == 0 && t_seq) {
        atomicAdd(&ctr->total_seq, t_seq); atomicAdd(&ctr->good_seq, g_seq);
        atomicAdd(&ctr->total_len, t_len); atomicAdd(&ctr->good_len, g_len);
        atomicAdd(&ctr->kmers, kmers);
    }
}

// ------------------------------------------------------------------------------------------
// table operations
// ------------------------------------------------------------------------------------------
// count += 1 for `key` (Long2ShortHashMap.addAndBound with inc = 1).  Thread-per-key linear
// probing; one 128-bit load brings key and count of a slot (same sector).  Returns true when
// this call claimed a new slot.
__device__ __forceinline__ bool table_upsert1_at(Slot *__restrict__ tab, uint64_t cap, uint64_t i, uint64_t key) {
    for (;;) {
        const ulonglong2 s = ld_cg_u64x2(&tab[i]);
        if (s.x == key) {
            // saturation: once the stored count reached 32767 further increments are no-ops
            if ((uint32_t)s.y < MAX_COUNT) atomicAdd(&tab[i].count, 1u);
            return false;
        }
        if (s.x == EMPTY_KEY) {
            const unsigned long long prev = atomicCAS(&tab[i].key, EMPTY_KEY, (unsigned long long)key);
            if (prev == EMPTY_KEY) { atomicAdd(&tab[i].count, 1u); return true; }
            if (prev == key) { atomicAdd(&tab[i].count, 1u); return false; }
        }
        if (++i == cap) i = 0;
    }
}
__device__ __forceinline__ bool table_upsert1(Slot *__restrict__ tab, uint64_t cap, uint64_t key) {
    return table_upsert1_at(tab, cap, home_slot(key, cap), key);
}

// count = sat_add(count, inc) for arbitrary inc (rehash, pre-aggregated (key,count) pairs).
__device__ __forceinline__ bool table_upsert_n_at(Slot *__restrict__ tab, uint64_t cap, uint64_t i, uint64_t key, uint32_t inc) {
    bool claimed = false;
    for (;;) {
        unsigned long long cur = ld_cg_u64x2(&tab[i]).x;
        if (cur == EMPTY_KEY) {
            cur = atomicCAS(&tab[i].key, EMPTY_KEY, (unsigned long long)key);
            if (cur == EMPTY_KEY) { claimed = true; cur = key; }
        }
        if (cur == key) {
            uint32_t old = *(volatile uint32_t *)&tab[i].count;
            for (;;) {
                if (old >= MAX_COUNT) break;
                uint32_t nv = old + inc; if (nv > MAX_COUNT || nv < old) nv = MAX_COUNT;
                const uint32_t seen = atomicCAS(&tab[i].count, old, nv);
                if (seen == old) break;
                old = seen;
            }
            return claimed;
        }
        if (++i == cap) i = 0;
    }
}
__device__ __forceinline__ bool table_upsert_n(Slot *__restrict__ tab, uint64_t cap, uint64_t key, uint32_t inc) {
    return table_upsert_n_at(tab, cap, home_slot(key, cap), key, inc);
}

// ------------------------------------------------------------------------------------------
// Windowed placement of the minimizer-placed tables.  A heavy minimizer (poly-A tails, low-complexity
// sequence; or plain skew when regions are small) can own more distinct k-mers than its region has
// slots, and linear probing that runs on into the neighbouring regions would build one huge cluster
// that every later insert walks.  So a key lives either
//   (1) in its PRIMARY window: the first `win` slots from its home slot, wrapping inside its region, or
//   (2) if that window held `win` other keys when it arrived, in the first of its SECONDARY windows
//       with room: `win` slots at home_slot(key ^ SALT * a), a = 1, 2, ... (a fresh, uniformly
//       distributed position per attempt, i.e. the free slots of arbitrary regions).
// Slots never become empty again, so the rule is stable: a key that went to (2) finds its earlier
// windows full for ever, and a key placed in a window is met before any empty slot of that window.
// win == 0 selects the plain rule (linear probing from the home slot, running on).
// ------------------------------------------------------------------------------------------
constexpr unsigned long long SECONDARY_SALT = 0x9ddfea08eb382d69ULL;
__device__ __forceinline__ uint64_t secondary_home(uint64_t key, uint64_t cap, uint32_t attempt) {
    return home_slot(key ^ (SECONDARY_SALT * (unsigned long long)attempt), cap);
}

// one slot of a probe sequence: returns true when the key was found or placed there
template <bool ONE>
__device__ __forceinline__ bool slot_try(Slot *__restrict__ tab, uint64_t i, uint64_t key, uint32_t inc, bool &claimed) {
    const ulonglong2 sl = ld_cg_u64x2(&tab[i]);
    unsigned long long cur = sl.x;
    bool mine = false;
    if (cur == EMPTY_KEY) {
        cur = atomicCAS(&tab[i].key, EMPTY_KEY, (unsigned long long)key);
        if (cur == EMPTY_KEY) { mine = true; cur = key; }
    }
    if (cur != key) return false;
    if (ONE) { if (mine || (uint32_t)sl.y < MAX_COUNT) atomicAdd(&tab[i].count, 1u); }
    else {
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

// Secondary positions: windows of `win` slots at home_slot(key ^ SALT * a), a = 1, 2, ... -- a fresh uniform position per
// attempt, so that a key never has to walk through a long run of full slots (an overfull region is one).
template <bool ONE>
__device__ __forceinline__ bool secondary_upsert(Slot *__restrict__ tab, uint64_t cap, uint32_t win, uint64_t key, uint32_t inc) {
    bool claimed = false;
    for (uint32_t a = 1;; a++) {
        uint64_t i = secondary_home(key, cap, a);
        for (uint32_t st = 0; st < win; st++) {
            if (slot_try<ONE>(tab, i, key, inc, claimed)) return claimed;
            if (++i == cap) i = 0;
        }
    }
}

template <bool ONE>
__device__ __forceinline__ bool placed_upsert_at(Slot *__restrict__ tab, uint64_t cap, int shift, uint32_t win,
                                                 uint64_t home, uint64_t key, uint32_t inc) {
    if (!win) return ONE ? table_upsert1_at(tab, cap, home, key) : table_upsert_n_at(tab, cap, home, key, inc);
    const uint64_t mask = (1ull << shift) - 1ull;
    const uint64_t rbase = home & ~mask;
    uint64_t off = home & mask;
    const uint32_t steps = (uint64_t)win < mask + 1 ? win : (uint32_t)(mask + 1);
    bool claimed = false;
    for (uint32_t st = 0; st < steps; st++, off = (off + 1) & mask)
        if (slot_try<ONE>(tab, rbase | off, key, inc, claimed)) return claimed;
    return secondary_upsert<ONE>(tab, cap, win, key, inc);
}

// ------------------------------------------------------------------------------------------
// Split (structure-of-arrays) table of the region-blocked variant: keys[cap] (8 B) and counts[cap]
// (4 B) in two arrays.  When every upsert hits L2, the 16-byte slot's "key and count share a
// sector" rule buys nothing, and measured on B200 an 8-byte key load + red.add on a different
// sector runs at 136 G upserts/s against 74-87 G/s for the 16-byte-slot pattern (mfkc_gups_ex
// modes 4 vs 3).  Increments are blind red.adds (no count load); counts are clamped to 32767 when
// read, and the host clamps the array between drains and keeps the k-mers of one drain below
// 4.0e9, so the u32 can never wrap.  The direct variant keeps the 16-byte slots (one DRAM sector).
// ------------------------------------------------------------------------------------------
struct TabSoA {
    unsigned long long *keys;
    uint32_t *counts;
    uint64_t cap;
};
__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ bool soa_upsert1_at(const TabSoA &t, uint64_t i, uint64_t key) {
    for (;;) {
        const unsigned long long cur = ld_cg_u64(&t.keys[i]);
        if (cur == key) { atomicAdd(&t.counts[i], 1u); return false; }
        if (cur == EMPTY_KEY) {
            const unsigned long long prev = atomicCAS(&t.keys[i], EMPTY_KEY, (unsigned long long)key);
            if (prev == EMPTY_KEY) { atomicAdd(&t.counts[i], 1u); return true; }
            if (prev == key) { atomicAdd(&t.counts[i], 1u); return false; }
        }
        if (++i == t.cap) i = 0;
    }
}
__device__ __forceinline__ bool soa_upsert_n_at(const TabSoA &t, uint64_t i, uint64_t key, uint32_t inc) {
    bool claimed = false;
    for (;;) {
        unsigned long long cur = ld_cg_u64(&t.keys[i]);
        if (cur == EMPTY_KEY) {
            cur = atomicCAS(&t.keys[i], EMPTY_KEY, (unsigned long long)key);
            if (cur == EMPTY_KEY) { claimed = true; cur = key; }
        }
        if (cur == key) {
            uint32_t old = *(volatile uint32_t *)&t.counts[i];
            for (;;) {
                if (old >= MAX_COUNT) break;
                uint32_t nv = old + inc; if (nv > MAX_COUNT || nv < old) nv = MAX_COUNT;
                const uint32_t seen = atomicCAS(&t.counts[i], old, nv);
                if (seen == old) break;
                old = seen;
            }
            return claimed;
        }
        if (++i == t.cap) i = 0;
    }
}
// counts[i] = min(counts[i], 32767): run between two drains (see above)
__global__ void __launch_bounds__(256)
soa_clamp_kernel(uint32_t *__restrict__ counts, uint64_t cap) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x)
        if (counts[i] > MAX_COUNT) counts[i] = MAX_COUNT;
}

// table accessors so that the super-k-mer kernels can be instantiated for either layout
struct TabAoS {
    Slot *tab; uint64_t cap;
    // shift / win: region size and primary window of the windowed placement (win = 0: plain rule), see placed_upsert_at
    __device__ __forceinline__ bool upsert1_at(uint64_t i, uint64_t key, int shift, uint32_t win) const { return placed_upsert_at<true>(tab, cap, shift, win, i, key, 1u); }
    __device__ __forceinline__ void prefetch_region(uint64_t first_slot, int shift, uint32_t t, uint32_t nt) const {
        const char *b = reinterpret_cast<const char *>(tab + first_slot);
        const uint64_t bytes = sizeof(Slot) << shift;
        for (uint64_t off = (uint64_t)t * 128; off < bytes; off += (uint64_t)nt * 128) prefetch_l2(b + off);
    }
};
struct TabSoAOps {
    TabSoA t;
    __device__ __forceinline__ bool upsert1_at(uint64_t i, uint64_t key, int, uint32_t) const { return soa_upsert1_at(t, i, key); }
    __device__ __forceinline__ void prefetch_region(uint64_t first_slot, int shift, uint32_t th, uint32_t nt) const {
        const char *kb = reinterpret_cast<const char *>(t.keys + first_slot);
        const char *cb = reinterpret_cast<const char *>(t.counts + first_slot);
        for (uint64_t off = (uint64_t)th * 128; off < (8ull << shift); off += (uint64_t)nt * 128) prefetch_l2(kb + off);
        for (uint64_t off = (uint64_t)th * 128; off < (4ull << shift); off += (uint64_t)nt * 128) prefetch_l2(cb + off);
    }
};

// ------------------------------------------------------------------------------------------
// Minimizer placement (region-blocked variant).  The table is n_regions x 2^region_shift slots;
// a k-mer lives in the region chosen by the smallest hash among its canonical m-mers
// (m = min(k, 12)) and, inside the region, at mix64(key) mod 2^region_shift (linear probing may
// run on into the next region).  Both strands of a k-mer have the same canonical m-mers, so the
// region is a function of the canonical key alone; consecutive k-mers of a read mostly share
// their minimizer, which is what lets the extraction kernel stage whole runs ("super-k-mers")
// per region instead of single keys.
// ------------------------------------------------------------------------------------------
constexpr int MINI_M = 12;
__host__ __device__ __forceinline__ int minimizer_len(int k) { return k < MINI_M ? k : MINI_M; }
// Minimizer length of the bin-local count (bincount.cuh), 12..16: the k-mer instances concentrate on the ~2/(k-m+2) of
// the m-mers with the smallest hashes, so a bin is the sum of (4^m / 2) * 0.1 / bins "effective" minimizers; with fewer
// than a few dozen of them per bin the bins' sizes scatter too much (measured: 12-mers, 447 k bins -> 5 % of the bins
// overflow a 2.6x segment).  4^m >= 640 * bins keeps >= 32 per bin.  (The table placement keeps m = 12.)
__host__ __device__ __forceinline__ int bin_minimizer_len(int k, uint64_t bins_total) {
    int m = MINI_M;
    while (m < 16 && (1ull << (2 * m)) < 640ull * bins_total) m++;
    return k < m ? k : m;
}
__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {       // MurmurHash3 fmix32
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
// order of the m-mers inside a k-mer: a bijective multiply-xorshift (3 instructions; the minimizer
// only needs a fixed pseudo-random ORDER, the spreading over regions / shards re-mixes with fmix32)
__host__ __device__ __forceinline__ uint32_t mmer_hash(uint32_t x) {
    x *= 0x9E3779B1u;
    return x ^ (x >> 15);
}
// spreading of the minimizer hash over the table regions (and, with its top bits, over the coarse buckets of the
// peer-memory staging).  The minimum of ~20 hashes is concentrated near 0, so it is re-mixed: multiply - xorshift -
// multiply, the high bits are the ones used (4 instructions; the extraction kernel evaluates it for all 16 start
// positions of a thread).  The owner shard keeps the full, independent fmix32.
__host__ __device__ __forceinline__ uint32_t region_hash(uint32_t mh) {
    uint32_t x = mh * 0x85ebca6bu;
    x ^= x >> 15;
    return x * 0xc2b2ae35u;
}
__host__ __device__ __forceinline__ uint32_t region_of_minhash(uint32_t mh, uint32_t n_regions) {
    return (uint32_t)(((uint64_t)region_hash(mh) * n_regions) >> 32);
}
__host__ __device__ __forceinline__ uint32_t owner_of_minhash(uint32_t mh, uint32_t n_shards) {
    return (uint32_t)(((uint64_t)hash32(mh ^ 0x1b873593u) * n_shards) >> 32);
}
// minimizer hash of a canonical key (slow path: rehash, keys that arrive without their read)
__host__ __device__ __forceinline__ uint32_t minhash_of_key(uint64_t key, int k) {
    const int m = minimizer_len(k);
    const uint32_t mask = m == 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1u);
    uint32_t best = 0xFFFFFFFFu;
    uint32_t fw = 0, rc = 0;
    for (int i = 0; i < k; i++) {                       // bases from the most significant pair down
        const uint32_t c = (uint32_t)(key >> (2 * (k - 1 - i))) & 3u;
        fw = ((fw << 2) | c) & mask;
        rc = (rc >> 2) | ((3u - c) << (2 * m - 2));
        if (i >= m - 1) { const uint32_t h = mmer_hash(fw < rc ? fw : rc); best = h < best ? h : best; }
    }
    return best;
}
__host__ __device__ __forceinline__ uint64_t mini_home(uint64_t key, uint32_t region, int region_shift) {
    return ((uint64_t)region << region_shift) | (mix64(key) & ((1ull << region_shift) - 1ull));
}

__global__ void __launch_bounds__(256)
table_clear_kernel(Slot *__restrict__ tab, uint64_t cap) {
    // 16-byte stores, fully coalesced
    const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint4 *>(tab)[i] = e;
}

// table geometry + placement rule, passed by value to kernels
struct TableGeom {
    uint64_t cap;
    uint32_t n_regions;
    int region_shift;
    int k;
    int minimizer;          // 1: minimizer placement, 0: plain hash placement
    uint32_t win;           // > 0: windowed placement (placed_upsert_at)
};
__device__ __forceinline__ uint64_t geom_home(const TableGeom &g, uint64_t key) {
    if (!g.minimizer) return home_slot(key, g.cap);
    return mini_home(key, region_of_minhash(minhash_of_key(key, g.k), g.n_regions), g.region_shift);
}

__global__ void __launch_bounds__(256)
rehash_kernel(const Slot *__restrict__ old_tab, uint64_t old_cap, Slot *__restrict__ new_tab, TableGeom g) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < old_cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 s = ld_nc_u128(&old_tab[i]);
        const unsigned long long key = ((unsigned long long)s.y << 32) | s.x;
        if (key != EMPTY_KEY) placed_upsert_at<false>(new_tab, g.cap, g.region_shift, g.minimizer ? g.win : 0u, geom_home(g, key), key, s.z < MAX_COUNT ? s.z : MAX_COUNT);
    }
}

// ------------------------------------------------------------------------------------------
// sinks of the extraction kernel
// ------------------------------------------------------------------------------------------
struct SinkTable {              // hash variant: upsert into the HBM-resident table
    Slot *tab; uint64_t cap; Counters *ctr;
    static constexpr bool kPrefetch = true;
    __device__ __forceinline__ void prefetch(uint64_t key) const { prefetch_l2(&tab[home_slot(key, cap)]); }
    __device__ __forceinline__ uint32_t put(uint64_t key) const { return table_upsert1(tab, cap, key) ? 1u : 0u; }
    __device__ __forceinline__ void finish(uint32_t local) const {
        // local = number of new slots claimed by this thread: warp-reduce -> one atomic per warp
        for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if (lane_id() == 0 && local) atomicAdd(&ctr->distinct, (unsigned long long)local);
    }
};

struct SinkPresence {           // features-calculator reads mode: if (contains(key)) acc += 1
    FcSlot *tab; uint64_t cap; const uint32_t *bloom; uint32_t bmask;
    static constexpr bool kPrefetch = true;
    __device__ __forceinline__ void prefetch(uint64_t key) const {
        const uint64_t mx = mix64(key);
        if (bmask) { const uint32_t b = (uint32_t)mx & bmask; if (!((__ldg(&bloom[b >> 5]) >> (b & 31u)) & 1u)) return; }
        prefetch_l2(&tab[mulhi64(mx, cap)]);
    }
    __device__ __forceinline__ uint32_t put(uint64_t key) const;
    __device__ __forceinline__ void finish(uint32_t) const {}
};

// NumUtils.addAndBound(long,long) ([itmo]/utils/NumUtils.java:27-32) with Java's wrapping arithmetic.
__device__ __forceinline__ long long add_and_bound64(long long v, long long inc) {
    const long long lim = (long long)(0x7fffffffffffffffULL - (unsigned long long)inc);
    if (v > lim) return 0x7fffffffffffffffLL;
    return (long long)((unsigned long long)v + (unsigned long long)inc);
}

// if (hm.contains(key)) hm.addAndBound(key, inc)   (src/io/IOUtils.java:583-587, 817-821)
// A one-hash Bloom bitmap in front (16 bits per component k-mer, at most 64 MiB: L2-resident): most records of a sample are
// not component k-mers, and the bitmap turns their random DRAM probe of the set into an L2 hit (no false negatives, so
// results are unchanged).  bmask = 0: no filter.
__device__ __forceinline__ void fc_accumulate(FcSlot *__restrict__ tab, uint64_t cap, uint64_t key, long long inc,
                                              const uint32_t *__restrict__ bloom = nullptr, uint32_t bmask = 0) {
    const uint64_t mx = mix64(key);
    if (bmask) {
        const uint32_t b = (uint32_t)mx & bmask;
        if (!((__ldg(&bloom[b >> 5]) >> (b & 31u)) & 1u)) return;
    }
    uint64_t i = mulhi64(mx, cap);
    for (;;) {
        const unsigned long long cur = ld_cg_u64x2(&tab[i]).x;
        if (cur == EMPTY_KEY) return;                 // not a component k-mer
        if (cur == key) {
            unsigned long long old = *(volatile unsigned long long *)&tab[i].acc;
            for (;;) {
                const unsigned long long nv = (unsigned long long)add_and_bound64((long long)old, inc);
                const unsigned long long seen = atomicCAS((unsigned long long *)&tab[i].acc, old, nv);
                if (seen == old) return;
                old = seen;
            }
        }
        if (++i == cap) i = 0;
    }
}
__device__ __forceinline__ uint32_t SinkPresence::put(uint64_t key) const { fc_accumulate(tab, cap, key, 1, bloom, bmask); return 0; }

// ------------------------------------------------------------------------------------------
// K1+K2(+K3): flat extraction.  The batch is one concatenated ASCII stream; thread t of a tile
// owns the 16 bases [16w, 16w+16) (one aligned 128-bit load), packs them to one 32-bit word in
// shared memory (base 0 in the top bit pair), and produces the (up to) 16 k-mers that START in
// its word from the 96-bit window (own word + two successors; 2 halo words per tile).  The
// forward k-mer is a funnel shift of the window, the reverse complement is rolled
// (ShortKmer.shiftRight, ShortKmer.java:68-71) from a bit-reversal seed (KmerUtils.java:12-22).
// ------------------------------------------------------------------------------------------
constexpr int EX_THREADS = 256;

__device__ __forceinline__ uint64_t revcomp64(uint64_t fw, int k) {
    // reverse the 2-bit groups of fw, complement, right-align (KmerUtils.reverseComplement)
    uint64_t x = __brevll(fw);                                         // reverses single bits
    x = ((x & 0x5555555555555555ULL) << 1) | ((x >> 1) & 0x5555555555555555ULL);   // restore pair order
    return (~x) >> (64 - 2 * k);
}

// Computes the canonical keys starting in this thread's word.  keys[j] valid iff bit j of the
// returned mask is set.
__device__ __forceinline__ uint32_t kmers_of_word(uint32_t w0, uint32_t w1, uint32_t w2, uint64_t flag_bits,
                                                  long long limit /* n_bases - k - 16*w */, int k,
                                                  uint64_t (&keys)[16]) {
    // flags over [p, p+k-2]; for k == 1 only "dead read" flags exist and the window is [p, p]
    const uint64_t span = (k > 1) ? ((1ULL << (k - 1)) - 1ULL) : 1ULL;
    const int rs = 64 - 2 * k;
    const int top = 2 * k - 2;
    uint32_t valid = 0;
    uint64_t rc = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t hi = j ? __funnelshift_l(w1, w0, 2 * j) : w0;
        const uint32_t lo = j ? __funnelshift_l(w2, w1, 2 * j) : w1;
        const uint64_t fw = (((uint64_t)hi << 32) | lo) >> rs;
        if (j == 0) rc = revcomp64(fw, k);
        else rc = (rc >> 2) | ((uint64_t)((~(uint32_t)fw) & 3u) << top);
        keys[j] = fw < rc ? fw : rc;                                     // ShortKmer.toLong
        const bool ok = ((flag_bits >> j) & span) == 0 && (long long)j <= limit;
        valid |= (ok ? 1u : 0u) << j;
    }
    return valid;
}

// Shared front end of every extraction kernel: K1 (ASCII -> 2-bit words in shared memory, 2 halo
// words, boundary flags) and the 48-base window + validity inputs of the calling thread.
struct TileWord {
    uint32_t w0, w1, w2;      // own word + two successors (base 0 in bit 31 of w0)
    uint64_t fbits;           // 64 boundary flags starting at the thread's first base
    long long limit;          // last start offset j with p + k <= n_bases
    bool active;              // the thread's word lies inside the batch
};
template <int NT>
__device__ __forceinline__ TileWord load_tile_word(const uint8_t *__restrict__ bases, uint64_t n_bases,
                                                   const uint32_t *__restrict__ flags, uint64_t tile, int k,
                                                   uint32_t *s_words /* NT+2 */, uint32_t *s_flags /* NT/2+2 */,
                                                   uint32_t &bad) {
    const uint32_t tid = threadIdx.x;
    const uint64_t n_flag_words = (n_bases + 31) >> 5;
    const uint64_t w_base = tile * NT;
#pragma unroll
    for (int rep = 0; rep < 2; rep++) {
        if (rep == 1 && tid >= 2) break;
        const uint32_t slot = rep ? NT + tid : tid;
        const uint64_t w = w_base + slot;
        uint32_t word = 0;
        const uint64_t b0 = w << 4;
        if (b0 + 16 <= n_bases) {
            const uint4 v = ld_nc_u128(bases + b0);
            word = pack16(v);
            bad |= bad4(v.x) | bad4(v.y) | bad4(v.z) | bad4(v.w);
        } else if (b0 < n_bases) {                      // ragged tail of the batch
            for (uint32_t j = 0; j < 16 && b0 + j < n_bases; j++) {
                const uint32_t c = bases[b0 + j];
                bad |= bad4(c | 0x41414100u);
                word |= pack4(c) >> 6 << (30 - 2 * j);
            }
        }
        s_words[slot] = word;
    }
    // boundary flags for positions [16*w_base, 16*w_base + 16*NT + 64)
    if (tid < NT / 2 + 2) {
        const uint64_t fw = (w_base >> 1) + tid;
        s_flags[tid] = fw < n_flag_words ? flags[fw] : 0u;
    }
    __syncthreads();
    TileWord t;
    const uint64_t w = w_base + tid;
    t.active = (w << 4) < n_bases;
    t.w0 = s_words[tid]; t.w1 = s_words[tid + 1]; t.w2 = s_words[tid + 2];
    const uint32_t f0 = s_flags[tid >> 1], f1 = s_flags[(tid >> 1) + 1], f2 = s_flags[(tid >> 1) + 2];
    t.fbits = (tid & 1) ? (((uint64_t)f0 >> 16) | ((uint64_t)f1 << 16) | ((uint64_t)f2 << 48))
                        : ((uint64_t)f0 | ((uint64_t)f1 << 32));
    t.limit = (long long)n_bases - k - (long long)(w << 4);
    return t;
}

template <class Sink>
__global__ void __launch_bounds__(EX_THREADS)
extract_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, const uint32_t *__restrict__ flags,
               int k, Sink sink, Counters *__restrict__ ctr) {
    __shared__ uint32_t s_words[EX_THREADS + 2];
    __shared__ uint32_t s_flags[EX_THREADS / 2 + 2];
    const uint64_t n_tiles = (((n_bases + 15) >> 4) + EX_THREADS - 1) / EX_THREADS;
    uint32_t claimed = 0, bad = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileWord t = load_tile_word<EX_THREADS>(bases, n_bases, flags, tile, k, s_words, s_flags, bad);
        if (t.active) {                                  // K2: canonical k-mers starting in my word
            uint64_t keys[16];
            const uint32_t valid = kmers_of_word(t.w0, t.w1, t.w2, t.fbits, t.limit, k, keys);
            if (Sink::kPrefetch) {                       // K3: sink
#pragma unroll
                for (int j = 0; j < 16; j++) if (valid >> j & 1) sink.prefetch(keys[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; j++) if (valid >> j & 1) claimed += sink.put(keys[j]);
        }
        __syncthreads();
    }
    sink.finish(claimed);
    if (__any_sync(0xffffffffu, bad != 0) && lane_id() == 0) atomicAdd(&ctr->bad_chars, 1ULL);
}

// Bulk-append flavour (sort variant and shard bucketing): the keys of a tile are written to
// `out` grouped by bucket (n_buckets = 1 for the sort variant).  Positions are claimed with one
// atomic per (warp, bucket): lanes agree on their bucket via match.any.
struct BucketSink {
    unsigned long long *out;            // bucket b occupies out[bucket_base[b] .. )
    const uint64_t *bucket_base;        // device, n_buckets entries (unused when count_only)
    unsigned long long *bucket_cursor;  // device, n_buckets entries (zeroed before launch)
    uint64_t out_cap;
    uint32_t n_buckets;
    int count_only;                     // pass 1: only bucket_cursor[b] += 1
};

template <int dummy = 0>
__global__ void __launch_bounds__(EX_THREADS)
extract_bucket_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, const uint32_t *__restrict__ flags,
                      int k, BucketSink sink, Counters *__restrict__ ctr) {
    __shared__ uint32_t s_words[EX_THREADS + 2];
    __shared__ uint32_t s_flags[EX_THREADS / 2 + 2];
    const uint32_t tid = threadIdx.x;
    const uint64_t n_tiles = (((n_bases + 15) >> 4) + EX_THREADS - 1) / EX_THREADS;
    uint32_t bad = 0;
    unsigned long long overflow = 0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileWord t = load_tile_word<EX_THREADS>(bases, n_bases, flags, tile, k, s_words, s_flags, bad);

        uint64_t keys[16];
        uint32_t valid = 0;
        if (t.active) {
            const uint32_t w0 = t.w0, w1 = t.w1, w2 = t.w2;
            const uint64_t fbits = t.fbits;
            const long long limit = t.limit;
            valid = kmers_of_word(w0, w1, w2, fbits, limit, k, keys);
        }
        // all 32 lanes take part in the warp-level position claims
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const bool ok = (valid >> j) & 1;
            const uint32_t b = (ok && sink.n_buckets > 1) ? owner_shard(keys[j], sink.n_buckets) : 0u;
            const uint32_t active = __ballot_sync(0xffffffffu, ok);
            if (!active) continue;
            uint32_t peers;
            if (sink.n_buckets > 1) peers = __match_any_sync(0xffffffffu, ok ? b : 0xFFFFFFFFu);
            else peers = active;
            if (ok) {
                const uint32_t rank = __popc(peers & lanemask_lt());
                const int leader = __ffs(peers) - 1;
                unsigned long long pos = 0;
                if ((int)lane_id() == leader) pos = atomicAdd(&sink.bucket_cursor[b], (unsigned long long)__popc(peers));
                pos = __shfl_sync(peers, pos, leader);
                if (!sink.count_only) {
                    const uint64_t at = sink.bucket_base[b] + pos + rank;
                    if (at < sink.out_cap) sink.out[at] = keys[j]; else overflow++;
                }
            }
        }
        __syncthreads();
    }
    if (overflow) atomicAdd(&ctr->overflow, overflow);
    if (__any_sync(0xffffffffu, bad != 0) && lane_id() == 0) atomicAdd(&ctr->bad_chars, 1ULL);
}

// ------------------------------------------------------------------------------------------
// Region-blocked counting (the default hash variant).
//
// A table upsert that misses L2 costs a random 32-byte DRAM sector read plus its write-back;
// measured on B200 that path saturates at ~15 G upserts/s, while the same upsert on an
// L2-resident window runs at 65-79 G/s (profiles/, mfkc_gups_ex modes 1 and 3).  So keys are
// first PARTITIONED by the table region their home slot falls in (region = 2^region_shift slots,
// 8 MiB by default) into a staging buffer -- streaming writes -- and later DRAINED region by
// region, so that all upserts of a region hit L2 and the region's sectors cross the HBM bus
// once in and once out, whatever the number of k-mer instances.
//   phase A  extract_partition_kernel / partition_keys_kernel   (per submitted batch)
//   phase B  drain_regions_kernel                               (when staging is full / at flush)
// A region segment that is full sends its keys straight to the table (slow but exact), so the
// staging buffer never needs exact sizing.
// ------------------------------------------------------------------------------------------
constexpr int MAX_REGIONS = 2048;          // single-key staging flavour (shared-memory histogram of the regions)
constexpr int MAX_REGIONS_SKM = 1 << 22;   // super-k-mer flavour (one cursor per region; 2^22 regions of 2^12 slots = 256 GiB)
constexpr int PT_THREADS = 512;

struct RegionStage {
    unsigned long long *keys;     // region r owns keys[r*seg_cap, (r+1)*seg_cap)
    unsigned int *cursor;         // per-region fill; values above seg_cap mean "the rest went direct"
    uint64_t seg_cap;             // < 2^31
    uint32_t n_regions;
    int region_shift;             // log2(slots per region); table capacity = n_regions << region_shift
};

// Stage the (up to 16) keys of one thread.  s_hist must be zeroed and the block synchronised
// before the call; contains two block-wide barriers.
template <int NT>
__device__ __forceinline__ uint32_t stage_keys_block(const uint64_t (&keys)[16], uint32_t valid, const RegionStage &rs,
                                                     Slot *__restrict__ tab, uint64_t cap, uint32_t *s_hist) {
    uint32_t rk[8];                                     // rank-in-tile of key j: 16 bits each (tile <= 8192 keys)
#pragma unroll
    for (int j = 0; j < 16; j++) {
        uint32_t rank = 0;
        if ((valid >> j) & 1) {
            const uint32_t region = (uint32_t)(home_slot(keys[j], cap) >> rs.region_shift);
            rank = atomicAdd(&s_hist[region], 1u);
        }
        if (j & 1) rk[j >> 1] |= rank << 16; else rk[j >> 1] = rank;
    }
    __syncthreads();
    // one global reservation per (tile, non-empty region)
    for (uint32_t r = threadIdx.x; r < rs.n_regions; r += NT) {
        const uint32_t c = s_hist[r];
        if (c) s_hist[r] = atomicAdd(&rs.cursor[r], c);
    }
    __syncthreads();
    uint32_t claimed = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if ((valid >> j) & 1) {
            const uint32_t region = (uint32_t)(home_slot(keys[j], cap) >> rs.region_shift);   // cheaper than 16 live registers
            const uint32_t rank = (j & 1) ? (rk[j >> 1] >> 16) : (rk[j >> 1] & 0xFFFFu);
            const uint64_t pos = (uint64_t)s_hist[region] + rank;
            if (pos < rs.seg_cap) rs.keys[(uint64_t)region * rs.seg_cap + pos] = keys[j];
            else claimed += table_upsert1(tab, cap, keys[j]) ? 1u : 0u;     // segment full: count directly
        }
    }
    return claimed;
}

__global__ void __launch_bounds__(PT_THREADS, 2)
extract_partition_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, const uint32_t *__restrict__ flags,
                         int k, RegionStage rs, Slot *__restrict__ tab, uint64_t cap, Counters *__restrict__ ctr) {
    __shared__ uint32_t s_words[PT_THREADS + 2];
    __shared__ uint32_t s_flags[PT_THREADS / 2 + 2];
    __shared__ uint32_t s_hist[MAX_REGIONS];
    const uint32_t tid = threadIdx.x;
    const uint64_t n_tiles = (((n_bases + 15) >> 4) + PT_THREADS - 1) / PT_THREADS;
    uint32_t claimed = 0, bad = 0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t r = tid; r < rs.n_regions; r += PT_THREADS) s_hist[r] = 0;
        const TileWord t = load_tile_word<PT_THREADS>(bases, n_bases, flags, tile, k, s_words, s_flags, bad);
        uint64_t keys[16];
        uint32_t valid = 0;
        if (t.active) {
            const uint32_t w0 = t.w0, w1 = t.w1, w2 = t.w2;
            const uint64_t fbits = t.fbits;
            const long long limit = t.limit;
            valid = kmers_of_word(w0, w1, w2, fbits, limit, k, keys);
        }
        claimed += stage_keys_block<PT_THREADS>(keys, valid, rs, tab, cap, s_hist);
        __syncthreads();
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
    if (__any_sync(0xffffffffu, bad != 0) && lane_id() == 0) atomicAdd(&ctr->bad_chars, 1ULL);
}

// Phase A, second flavour: every key reserves its staging position with one returning atomic on
// the region cursor in L2 (no shared-memory histogram, no block barriers).  The 16 atomics of a
// thread are issued back to back, the dependent stores follow, so ~16 requests per thread are in
// flight.  Region cursors are few (<= 2048) but live in L2, whose atomic units sustain ~190 G
// atomics/s on resident lines (mfkc_gups_ex mode 0).
__global__ void __launch_bounds__(EX_THREADS)
extract_stage_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, const uint32_t *__restrict__ flags,
                     int k, RegionStage rs, Slot *__restrict__ tab, uint64_t cap, Counters *__restrict__ ctr) {
    __shared__ uint32_t s_words[EX_THREADS + 2];
    __shared__ uint32_t s_flags[EX_THREADS / 2 + 2];
    const uint32_t tid = threadIdx.x;
    const uint64_t n_tiles = (((n_bases + 15) >> 4) + EX_THREADS - 1) / EX_THREADS;
    uint32_t claimed = 0, bad = 0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileWord t = load_tile_word<EX_THREADS>(bases, n_bases, flags, tile, k, s_words, s_flags, bad);
        if (t.active) {
            const uint32_t w0 = t.w0, w1 = t.w1, w2 = t.w2;
            const uint64_t fbits = t.fbits;
            uint64_t keys[16];
            const long long limit = t.limit;
            const uint32_t valid = kmers_of_word(w0, w1, w2, fbits, limit, k, keys);
            uint32_t pos[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                pos[j] = 0;
                if ((valid >> j) & 1) {
                    const uint32_t region = (uint32_t)(home_slot(keys[j], cap) >> rs.region_shift);
                    pos[j] = atomicAdd(&rs.cursor[region], 1u);
                }
            }
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if ((valid >> j) & 1) {
                    const uint32_t region = (uint32_t)(home_slot(keys[j], cap) >> rs.region_shift);
                    if (pos[j] < rs.seg_cap) rs.keys[(uint64_t)region * rs.seg_cap + pos[j]] = keys[j];
                    else claimed += table_upsert1(tab, cap, keys[j]) ? 1u : 0u;   // segment full: count directly
                }
            }
        }
        __syncthreads();
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
    if (__any_sync(0xffffffffu, bad != 0) && lane_id() == 0) atomicAdd(&ctr->bad_chars, 1ULL);
}

// phase A for keys that already exist as an array (receive side of the shard exchange)
__global__ void __launch_bounds__(PT_THREADS, 2)
partition_keys_kernel(const unsigned long long *__restrict__ in, uint64_t n, RegionStage rs,
                      Slot *__restrict__ tab, uint64_t cap, Counters *__restrict__ ctr) {
    __shared__ uint32_t s_hist[MAX_REGIONS];
    const uint32_t tid = threadIdx.x;
    const uint64_t tile_keys = (uint64_t)PT_THREADS * 16;
    const uint64_t n_tiles = (n + tile_keys - 1) / tile_keys;
    uint32_t claimed = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t r = tid; r < rs.n_regions; r += PT_THREADS) s_hist[r] = 0;
        __syncthreads();
        uint64_t keys[16];
        uint32_t valid = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {                   // coalesced: consecutive threads, consecutive keys
            const uint64_t i = tile * tile_keys + (uint64_t)j * PT_THREADS + tid;
            keys[j] = 0;
            if (i < n) { keys[j] = in[i]; valid |= 1u << j; }
        }
        claimed += stage_keys_block<PT_THREADS>(keys, valid, rs, tab, cap, s_hist);
        __syncthreads();
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

// phase B: blocks_per_region consecutive CTAs own one region; the hardware launches CTAs in
// index order, so only (resident CTAs / blocks_per_region) regions are live at any time.
__global__ void __launch_bounds__(256)
drain_regions_kernel(RegionStage rs, uint32_t blocks_per_region, Slot *__restrict__ tab, uint64_t cap,
                     Counters *__restrict__ ctr) {
    const uint32_t region = blockIdx.x / blocks_per_region;
    const uint32_t sub = blockIdx.x % blocks_per_region;
    uint64_t n = rs.cursor[region];
    if (n > rs.seg_cap) n = rs.seg_cap;
    const unsigned long long *__restrict__ keys = rs.keys + (uint64_t)region * rs.seg_cap;
    uint32_t claimed = 0;
    constexpr int U = 4;                                  // independent upserts in flight per thread
    const uint64_t stride = (uint64_t)blocks_per_region * 256;
    for (uint64_t i0 = (uint64_t)sub * 256 + threadIdx.x; i0 < n; i0 += stride * U) {
        unsigned long long key[U]; uint64_t slot[U]; ulonglong2 s[U]; bool live[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t i = i0 + (uint64_t)u * stride;
            live[u] = i < n;
            key[u] = live[u] ? keys[i] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < U; u++) { slot[u] = home_slot(key[u], cap); if (live[u]) s[u] = ld_cg_u64x2(&tab[slot[u]]); }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!live[u]) continue;
            uint64_t i = slot[u];
            ulonglong2 v = s[u];
            for (;;) {                                    // same protocol as table_upsert1, first probe preloaded
                if (v.x == key[u]) { if ((uint32_t)v.y < MAX_COUNT) atomicAdd(&tab[i].count, 1u); break; }
                if (v.x == EMPTY_KEY) {
                    const unsigned long long prev = atomicCAS(&tab[i].key, EMPTY_KEY, key[u]);
                    if (prev == EMPTY_KEY) { atomicAdd(&tab[i].count, 1u); claimed++; break; }
                    if (prev == key[u]) { atomicAdd(&tab[i].count, 1u); break; }
                }
                if (++i == cap) i = 0;
                v = ld_cg_u64x2(&tab[i]);
            }
        }
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

// ------------------------------------------------------------------------------------------
// Super-k-mer staging (default flavour of the region-blocked variant).
// Phase A: per thread, the 16 k-mers starting in its word are cut into runs of consecutive valid
// k-mers whose minimizers fall in the same table region; each run becomes ONE 16-byte record
//   x,y,z = the run's bases, 2 bits each, first base in bit 31 of x (len + k - 1 <= 46 bases);
//           the low 4 bits of z hold len - 1 (len = 1..16 k-mers)
//   w     = the minimizer hash (region / owner shard derive from it)
// appended to the region's segment (one returning atomic on the region cursor per RECORD, i.e. per
// ~6 k-mers, instead of per key).  Phase B expands the records again next to the table region.
// Staging traffic drops from 8 B to ~2.7 B per k-mer instance and the scatter work by the run length.
// ------------------------------------------------------------------------------------------
struct SkmStage {
    uint4 *recs;                  // region r owns recs[r*seg_cap, (r+1)*seg_cap)
    unsigned int *cursor;
    uint64_t seg_cap;             // records per region, < 2^31
    uint32_t n_regions;
    int region_shift;
    uint32_t win;                 // primary window of the windowed placement (0 = plain linear probing), see placed_upsert_at
    // MODE 3 / 4 (bin-local counting): records that find their bin's segment full are appended to this list together with
    // (segment, position past the segment); ovf_place_kernel files them into the chunk pool before the count
    uint4 *ovf; uint2 *ovf_meta; unsigned int *ovf_used; uint32_t ovf_cap;
    int mlen;                     // MODE 3 / 4: minimizer length (bin_minimizer_len); other modes use minimizer_len(k)
};

__device__ __forceinline__ uint32_t kmers_of_word(uint32_t w0, uint32_t w1, uint32_t w2, uint64_t flag_bits,
                                                  long long limit, int k, uint64_t (&keys)[16]);

// Overflow pool of the bin staging: the records of a bin beyond its segment go to chunks of 32 records; chunk j of segment
// `seg` lives wherever the probe sequence of (seg, j) first met a free or matching tag.  Writers claim (CLAIM = true, lock
// free: the first record of a chunk to arrive takes the slot, nobody ever waits), readers only look up.  Returns the chunk
// index or 0xFFFFFFFF (pool full / chunk never written).
constexpr unsigned long long OVF_EMPTY_TAG = ~0ull;
constexpr uint32_t OVF_NO_CHUNK = 0xFFFFFFFFu;
constexpr uint32_t OVF_MAX_PROBE = 4096;
template <bool CLAIM>
__device__ __forceinline__ uint32_t ovf_chunk_of(unsigned long long *tags, uint32_t n_chunks, uint32_t seg, uint32_t j, unsigned int *used) {
    const unsigned long long tag = ((unsigned long long)seg << 32) | j;
    uint32_t c = (uint32_t)(((uint64_t)hash32(seg * 0x9E3779B1u + j * 0x85EBCA6Bu + 0x27d4eb2fu) * n_chunks) >> 32);
    for (uint32_t probe = 0; probe < OVF_MAX_PROBE && probe < n_chunks; probe++) {
        unsigned long long cur = ld_cg_u64(&tags[c]);
        if (cur == tag) return c;
        if (cur == OVF_EMPTY_TAG) {
            if (!CLAIM) return OVF_NO_CHUNK;
            cur = atomicCAS(&tags[c], OVF_EMPTY_TAG, tag);
            if (cur == OVF_EMPTY_TAG) { if (used) atomicAdd(used, 1u); return c; }
            if (cur == tag) return c;
        }
        if (++c == n_chunks) c = 0;
    }
    return OVF_NO_CHUNK;
}
// Overflow list -> chunk pool (run before every count; filing a record twice is harmless).  The list is filled by the
// extraction kernel with one cursor atomic per record; doing the chunk lookup there instead costs that kernel registers.
__global__ void __launch_bounds__(256)
ovf_place_kernel(const uint4 *__restrict__ list, const uint2 *__restrict__ meta, const unsigned int *__restrict__ n_list, uint32_t list_cap,
                 uint4 *__restrict__ pool, unsigned long long *__restrict__ tags, uint32_t n_chunks, Counters *__restrict__ ctr) {
    uint32_t n = *n_list;
    if (n > list_cap) n = list_cap;
    uint32_t dropped = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 mt = meta[i];
        const uint32_t ch = ovf_chunk_of<true>(tags, n_chunks, mt.x, mt.y >> 5, nullptr);
        if (ch == OVF_NO_CHUNK) { dropped++; continue; }
        pool[(uint64_t)ch * 32 + (mt.y & 31u)] = list[i];
    }
    if (dropped) atomicAdd(&ctr->overflow, (unsigned long long)dropped);
}

// a record that found its region segment full: count its k-mers straight into the table
template <class Tab>
__device__ __noinline__ uint32_t skm_count_direct(uint4 rec, uint32_t region, int region_shift, uint32_t win, int k, Tab tb) {
    // rare path: rolled loop, no register arrays (its register need adds to the calling kernel's)
    const uint32_t len = (rec.z & 15u) + 1u;
    const uint32_t w2 = rec.z & ~15u;
    const int rs = 64 - 2 * k;
    uint32_t claimed = 0;
#pragma unroll 1
    for (uint32_t t = 0; t < len; t++) {
        const uint32_t h32 = __funnelshift_l(rec.y, rec.x, 2 * t);
        const uint32_t l32 = __funnelshift_l(w2, rec.y, 2 * t);
        const uint64_t fw = (((uint64_t)h32 << 32) | l32) >> rs;
        const uint64_t rc = revcomp64(fw, k);
        const uint64_t key = fw < rc ? fw : rc;
        claimed += tb.upsert1_at(mini_home(key, region, region_shift), key, region_shift, win) ? 1u : 0u;
    }
    return claimed;
}

// minimizer hash of each of the 16 k-mers that start in word 0 of the 48-base window.
// h[q] = hash of the canonical m-mer at base offset q; window of k-mer j = h[j .. j+w-1], w = k-m+1.
// For w >= 16 (k >= 27) every window contains h[j..15] and h[16..w-1], so
//   min(window j) = min(suffix[j], middle, prefix[j])   (van Herk / Gil-Werman split at 15|16)
// costs ~70 min operations instead of 16 x w; the split needs w at compile time, hence the dispatch.
template <int W>
__device__ __forceinline__ void window_mins_static(const uint32_t (&h)[36], uint32_t (&mh)[16]) {
    uint32_t suf[16];
    uint32_t run = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 15; j >= 0; j--) { run = min(run, h[j]); suf[j] = run; }
    uint32_t mid = 0xFFFFFFFFu;
#pragma unroll
    for (int q = 16; q < W; q++) mid = min(mid, h[q]);
    run = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (j) run = min(run, h[W - 1 + j]);
        mh[j] = min(min(suf[j], mid), run);
    }
}

// hashes of the canonical m-mers at base offsets q0 .. q0 + N - 1 of the window (w0, w1, w2): forward m-mer by funnel shift,
// reverse complement seeded once and rolled
template <int N>
__device__ __forceinline__ void mmer_hashes(uint32_t w0, uint32_t w1, uint32_t w2, int m, uint32_t (&h)[N]) {
    const int rs = 32 - 2 * m;
    const int top = 2 * m - 2;
    uint32_t rc = 0;
#pragma unroll
    for (int q = 0; q < N; q++) {
        const uint32_t lo_w = q < 16 ? w0 : (q < 32 ? w1 : w2);
        const uint32_t hi_w = q < 16 ? w1 : (q < 32 ? w2 : 0u);
        const int sh = 2 * (q & 15);
        const uint32_t v = sh ? __funnelshift_l(hi_w, lo_w, sh) : lo_w;      // 16 bases starting at offset q
        const uint32_t fw = v >> rs;
        if (q == 0) {
            uint32_t x = __brev(fw);
            x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
            rc = (~x) >> rs;
        } else {
            rc = (rc >> 2) | (((~fw) & 3u) << top);
        }
        h[q] = mmer_hash(fw < rc ? fw : rc);
    }
}

// minimizer hash of the 16 k-mers starting at offsets 0..15, from the m-mer hashes at offsets 0..35 (w = k - m + 1 per k-mer)
__device__ __forceinline__ void window_mins(const uint32_t (&h)[36], int w, uint32_t (&mh)[16]) {
    switch (w) {                                          // uniform branch
        case 20: window_mins_static<20>(h, mh); return;   // k = 31, m = 12
        case 19: window_mins_static<19>(h, mh); return;
        case 18: window_mins_static<18>(h, mh); return;
        case 17: window_mins_static<17>(h, mh); return;
        case 16: window_mins_static<16>(h, mh); return;
        default: break;
    }
#pragma unroll
    for (int j = 0; j < 16; j++) {
        uint32_t best = 0xFFFFFFFFu;
#pragma unroll
        for (int q = 0; q < 15; q++) if (q < w) best = min(best, h[j + q]);
        mh[j] = best;
    }
}

__device__ __forceinline__ void minhash_of_word(uint32_t w0, uint32_t w1, uint32_t w2, int k, int m, uint32_t (&mh)[16]) {
    uint32_t h[36];                                       // m-mer hashes at base offsets 0..35
    mmer_hashes<36>(w0, w1, w2, m, h);
    window_mins(h, k - m + 1, mh);
}

// MODE 0: bucket = table region of this GPU (single-GPU path, full segments fall back to direct upserts);
// MODE 1: bucket = owner shard (send buffer of the NCCL exchange; st.n_regions = number of shards,
//         kmer_count[b] receives the k-mer instances sent to shard b; a full segment is reported
//         through the cursor overshoot and the caller retries with a smaller batch);
// MODE 2: bucket = owner shard x coarse bucket of the minimizer hash (peer-memory exchange: the owner
//         drains these segments straight from this GPU's memory, see drain_p2p_kernel).
//         st.n_regions = number of shards, st.region_shift = log2(coarse buckets per shard);
//         kmer_count[owner] as in mode 1 (warp-aggregated: 8 hot addresses would serialise in L2).
// MODE 3: bucket = minimizer BIN of the bin-local count (bincount.cuh); st.n_regions = number of bins.  A full segment
//         sends the record to the overflow list (st.ovf); only when that is full too the record is dropped and reported.
// MODE 4: like 3 with shards: bucket = owner * bins_per_shard + bin (st.n_regions = shards, st.region_shift unused,
//         st.win = bins per shard); kmer_count[owner] as in mode 2.
__host__ __device__ __forceinline__ uint32_t coarse_of_minhash(uint32_t mh, int log2_buckets) {
    return log2_buckets ? region_hash(mh) >> (32 - log2_buckets) : 0u;              // monotone in the region index
}
// modes 3 / 4 (mask-based run loop, no table fallback inside the kernel) fit 64 registers: 4 CTAs per SM hide the cursor atomics
template <int MODE, class Tab>
__global__ void __launch_bounds__(EX_THREADS, (MODE >= 3 ? 4 : 2))
extract_skm_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, const uint32_t *__restrict__ flags,
                   int k, SkmStage st, Tab tb, Counters *__restrict__ ctr,
                   unsigned long long *__restrict__ kmer_count) {
    __shared__ uint32_t s_words[EX_THREADS + 2];
    __shared__ uint32_t s_flags[EX_THREADS / 2 + 2];
    __shared__ uint32_t s_rkey[(MODE == 3 || MODE == 4) ? 16 : 1][EX_THREADS];      // minimizer hashes of the 16 start positions of every thread
    __shared__ uint32_t s_h[(MODE == 3 || MODE == 4) ? 16 : 1][EX_THREADS + 2];     // m-mer hashes of the tile: [offset in word][word]
    const uint32_t tid = threadIdx.x;
    const uint64_t n_tiles = (((n_bases + 15) >> 4) + EX_THREADS - 1) / EX_THREADS;
    uint32_t claimed = 0, bad = 0, dropped = 0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileWord t = load_tile_word<EX_THREADS>(bases, n_bases, flags, tile, k, s_words, s_flags, bad);
        uint32_t hq[16];
        if constexpr (MODE == 3 || MODE == 4) {
            // every thread hashes the 16 m-mers that start in ITS word (instead of all 35 its k-mers touch) and shares them
            // through shared memory; threads 0 and 1 also do the two halo words of the tile
            mmer_hashes<16>(t.w0, t.w1, 0u, st.mlen, hq);
#pragma unroll
            for (int q = 0; q < 16; q++) s_h[q][tid] = hq[q];
            if (tid < 2) {
                uint32_t hh[16];
                mmer_hashes<16>(s_words[EX_THREADS + tid], tid == 0 ? s_words[EX_THREADS + 1] : 0u, 0u, st.mlen, hh);
#pragma unroll
                for (int q = 0; q < 16; q++) s_h[q][EX_THREADS + tid] = hh[q];
            }
            __syncthreads();
        }
        unsigned long long km_lo = 0, km_hi = 0;                // MODE 2: k-mers per owner of this tile, 16-bit fields
        if (t.active) {
            const uint32_t w0 = t.w0, w1 = t.w1, w2 = t.w2;
            const uint64_t fbits = t.fbits;
            // validity of the 16 start positions (same rule as kmers_of_word: no boundary flag in [j, j + k - 2] and
            // j <= limit), computed for all positions at once: smear the flags down by the window length with
            // doubling shifts (bit j of `inv` = OR of the flags j .. j + span_len - 1), ~45 instructions instead of 130
            const long long limit = t.limit;
            uint32_t valid;
            {
                const int span_len = k > 1 ? k - 1 : 1;
                uint64_t inv = fbits;
                for (int have = 1; have < span_len;) { const int step = have < span_len - have ? have : span_len - have; inv |= inv >> step; have += step; }
                const uint32_t lim_mask = limit >= 15 ? 0xFFFFu : (limit < 0 ? 0u : ((2u << (int)limit) - 1u));
                valid = ~(uint32_t)inv & lim_mask;
            }
            if constexpr (MODE == 3 || MODE == 4) { if (valid) {
                // Bin staging.  The runs of a thread (consecutive valid k-mers of one minimizer; 2.5 per thread on average,
                // 16 at most) are found with bit masks and walked in a loop: the minimizer hashes wait in shared memory (a
                // register array cannot be indexed by the run's start), the bin is derived per RUN, the cursor atomic of the
                // NEXT run is in flight while the record of the current one is built and stored.  (The fully unrolled
                // two-pass state machine below spent half of the kernel's instructions on its 2 x 17 predicated steps.)
                uint32_t mh[16];
                {
                    // m-mer hashes 0..15 are this thread's own (computed for the whole tile above), 16..34 its neighbours'
                    uint32_t h[36];
#pragma unroll
                    for (int q = 0; q < 16; q++) { h[q] = hq[q]; h[16 + q] = s_h[q][tid + 1]; }
#pragma unroll
                    for (int q = 0; q < 4; q++) h[32 + q] = q < 3 ? s_h[q][tid + 2] : 0xFFFFFFFFu;
                    window_mins(h, k - st.mlen + 1, mh);
                }
                uint32_t eq = 0;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    s_rkey[j][tid] = mh[j];
                    if (j && mh[j] == mh[j - 1]) eq |= 1u << j;
                }
                // MODE 4: owner in the top byte (<= 16 shards), bin of the shard below (< 2^24)
                auto key_of = [&](uint32_t mhv) {
                    return MODE == 4 ? ((owner_of_minhash(mhv, st.n_regions) << 24) | region_of_minhash(mhv, st.win)) : region_of_minhash(mhv, st.n_regions);
                };
                const uint32_t starts = valid & ~((valid << 1) & eq);       // valid, and not the continuation of the k-mer before
                const uint32_t stops = (~valid | starts) | (1u << 16);       // a run ends in front of the next start / invalid position
                const uint32_t seg32 = (uint32_t)st.seg_cap;
                auto emit_run = [&](uint32_t sj, uint32_t key, uint32_t bucket, uint32_t p) {
                    const uint32_t len = (uint32_t)__ffs(stops & ~((2u << sj) - 1u)) - 1u - sj;
                    const uint32_t sh = 2u * sj;                          // normalise: first base of the run -> base 0
                    uint4 rec;
                    rec.x = __funnelshift_l(w1, w0, sh);
                    rec.y = __funnelshift_l(w2, w1, sh);
                    rec.z = ((w2 << sh) & ~15u) | (len - 1u);
                    rec.w = key;
                    if (p < seg32) st.recs[(uint64_t)bucket * seg32 + p] = rec;
                    else {                                               // segment full: overflow list (filed into the chunk pool before the count)
                        const uint32_t o = atomicAdd(st.ovf_used, 1u);
                        if (o < st.ovf_cap) { st.ovf[o] = rec; st.ovf_meta[o] = make_uint2(bucket, p - seg32); } else dropped++;
                    }
                    if (MODE == 4) {                                     // k-mers per owner (counted even if dropped: mfkc_flush reports the drop)
                        const uint32_t ow = key >> 24;
                        if (st.n_regions <= 8) { if (ow < 4) km_lo += (unsigned long long)len << (16 * ow); else km_hi += (unsigned long long)len << (16 * (ow - 4)); }
                        else atomicAdd(&kmer_count[ow], (unsigned long long)len);
                    }
                };
                // The first EX_RUNS runs of the thread (practically always all of them): every cursor atomic is issued before the
                // first record is stored, so their ~1 us round trips overlap (ncu of the one-ahead loop: long scoreboard 9 of
                // 20 stall cycles per issue).  Positions and keys stay in registers: the loops are unrolled and predicated.
                constexpr int EX_RUNS = 6;
                uint32_t left = starts;
                uint32_t r_pos[EX_RUNS], r_key[EX_RUNS], r_sj[EX_RUNS];
#pragma unroll
                for (int r = 0; r < EX_RUNS; r++) {
                    r_pos[r] = 0; r_key[r] = 0; r_sj[r] = 0xFFu;
                    if (left) {
                        const uint32_t sj = __ffs(left) - 1;
                        left &= left - 1;
                        const uint32_t key = key_of(s_rkey[sj][tid]);
                        const uint32_t bucket = MODE == 4 ? (key >> 24) * st.win + (key & 0xFFFFFFu) : key;
                        r_sj[r] = sj; r_key[r] = key;
                        r_pos[r] = atomicAdd(&st.cursor[bucket], 1u);
                    }
                }
#pragma unroll
                for (int r = 0; r < EX_RUNS; r++)
                    if (r_sj[r] != 0xFFu) {
                        const uint32_t key = r_key[r];
                        emit_run(r_sj[r], key, MODE == 4 ? (key >> 24) * st.win + (key & 0xFFFFFFu) : key, r_pos[r]);
                    }
                while (left) {                                           // more than EX_RUNS runs in 16 positions: one by one
                    const uint32_t sj = __ffs(left) - 1;
                    left &= left - 1;
                    const uint32_t key = key_of(s_rkey[sj][tid]);
                    const uint32_t bucket = MODE == 4 ? (key >> 24) * st.win + (key & 0xFFFFFFu) : key;
                    emit_run(sj, key, bucket, atomicAdd(&st.cursor[bucket], 1u));
                }
            } } else if (valid) {
                uint32_t mh[16];
                minhash_of_word(w0, w1, w2, k, minimizer_len(k), mh);
                // Cut into runs of consecutive valid k-mers (fully unrolled: every register array keeps
                // compile-time indices).  Local staging: a run = same table REGION.  Send buffer
                // (BY_OWNER): a run = same MINIMIZER HASH, because the receiver derives the region of
                // the whole record from that one hash, under a table geometry the sender does not know.
                // Two passes over the same state machine: pass 1 issues every cursor atomic of the
                // thread back to back (their ~1 us round trips overlap), pass 2 writes the records.
                constexpr bool BY_OWNER = MODE == 1 || MODE == 2;
                auto bucket_of = [&](uint32_t mhv, uint32_t key) {
                    if (MODE == 0 || MODE == 3 || MODE == 4) return key;
                    const uint32_t o = owner_of_minhash(mhv, st.n_regions);
                    return MODE == 1 ? o : ((o << st.region_shift) | coarse_of_minhash(mhv, st.region_shift));
                };
                const uint32_t seg32 = (uint32_t)st.seg_cap;
                uint32_t rkey[16];                                  // run key of every start position, computed once
#pragma unroll
                for (int j = 0; j < 16; j++)
                    rkey[j] = BY_OWNER ? mh[j]
                            : (MODE == 4 ? owner_of_minhash(mh[j], st.n_regions) * st.win + region_of_minhash(mh[j], st.win)
                                         : region_of_minhash(mh[j], st.n_regions));
                uint32_t pos[17];
                {
                    uint32_t run_key = 0, run_mh = 0;
                    bool in_run = false;
#pragma unroll
                    for (int j = 0; j <= 16; j++) {
                        const bool v = j < 16 && ((valid >> j) & 1);
                        const uint32_t mhj = mh[j < 16 ? j : 15];
                        const uint32_t key = rkey[j < 16 ? j : 15];
                        pos[j] = 0;
                        if (in_run && (!v || key != run_key)) { pos[j] = atomicAdd(&st.cursor[bucket_of(run_mh, run_key)], 1u); in_run = false; }
                        if (v && !in_run) { in_run = true; run_key = key; run_mh = mhj; }
                    }
                }
                {
                    uint32_t run_start = 0, run_key = 0, run_mh = 0;
                    bool in_run = false;
#pragma unroll
                    for (int j = 0; j <= 16; j++) {
                        const bool v = j < 16 && ((valid >> j) & 1);
                        const uint32_t mhj = mh[j < 16 ? j : 15];
                        const uint32_t key = rkey[j < 16 ? j : 15];
                        if (in_run && (!v || key != run_key)) {
                            const uint32_t len = (uint32_t)j - run_start;
                            const int sh = 2 * (int)run_start;      // normalise: first base of the run -> base 0
                            uint4 rec;
                            rec.x = sh ? __funnelshift_l(w1, w0, sh) : w0;
                            rec.y = sh ? __funnelshift_l(w2, w1, sh) : w1;
                            rec.z = ((sh ? (w2 << sh) : w2) & ~15u) | (len - 1);
                            rec.w = run_mh;
                            const uint32_t bucket = bucket_of(run_mh, run_key);
                            if (pos[j] < seg32) {
                                // local staging holds < 2^32 records (reserve_staging caps it): 32-bit index arithmetic
                                if (MODE == 0) st.recs[bucket * seg32 + pos[j]] = rec; else st.recs[(uint64_t)bucket * seg32 + pos[j]] = rec;   // MODE 0: < 2^32 records
                                if (MODE == 1) atomicAdd(&kmer_count[bucket], (unsigned long long)len);
                                if (MODE == 2 || MODE == 4) {
                                    const uint32_t o = MODE == 4 ? bucket / st.win : bucket >> st.region_shift;
                                    if (st.n_regions <= 8) { if (o < 4) km_lo += (unsigned long long)len << (16 * o); else km_hi += (unsigned long long)len << (16 * (o - 4)); }
                                    else atomicAdd(&kmer_count[o], (unsigned long long)len);
                                }
                            } else if (MODE == 3 || MODE == 4) {    // segment full: overflow list (filed into the chunk pool before the count)
                                const uint32_t o = atomicAdd(st.ovf_used, 1u);
                                if (o < st.ovf_cap) { st.ovf[o] = rec; st.ovf_meta[o] = make_uint2(bucket, pos[j] - seg32); } else dropped++;
                                if (MODE == 4) {                    // (counted for its owner even if dropped: mfkc_flush reports the drop)
                                    const uint32_t ow = bucket / st.win;
                                    if (st.n_regions <= 8) { if (ow < 4) km_lo += (unsigned long long)len << (16 * ow); else km_hi += (unsigned long long)len << (16 * (ow - 4)); }
                                    else atomicAdd(&kmer_count[ow], (unsigned long long)len);
                                }
                            } else if (!BY_OWNER) {                 // segment full: count the run directly (slow, exact)
                                claimed += skm_count_direct(rec, bucket, st.region_shift, st.win, k, tb);
                            } else if (MODE == 2) dropped++;        // reported as an error by mfkc_flush (segments are sized with 2x slack)
                            in_run = false;
                        }
                        if (v && !in_run) { in_run = true; run_start = (uint32_t)j; run_key = key; run_mh = mhj; }
                    }
                }
            }
        }
        if ((MODE == 2 || MODE == 4) && st.n_regions <= 8) {    // one atomic per (warp, owner) and tile
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
    if ((MODE == 2 || MODE == 3 || MODE == 4) && dropped) atomicAdd(&ctr->overflow, (unsigned long long)dropped);
}

// Send side for up to 8 owner shards with block-level aggregation.  With only G <= 8 destination
// cursors, one returning atomic per record serialises on G addresses in L2 (measured: 46 ms for
// 4.8e8 k-mers on 2 shards).  Here a tile counts its records per owner first (packed 16-bit fields,
// warp scan + shared prefix), reserves ONE range per (tile, owner) and then writes the records:
// ~90x fewer global atomics, and each owner's records of a tile leave as one contiguous burst.
__device__ __forceinline__ void pk_add(unsigned long long &lo, unsigned long long &hi, uint32_t o, uint32_t v) {
    if (o < 4) lo += (unsigned long long)v << (16 * o); else hi += (unsigned long long)v << (16 * (o - 4));
}
__device__ __forceinline__ uint32_t pk_get(unsigned long long lo, unsigned long long hi, uint32_t o) {
    return (uint32_t)((o < 4 ? lo >> (16 * o) : hi >> (16 * (o - 4))) & 0xFFFFu);
}

__global__ void __launch_bounds__(EX_THREADS)
extract_skm_owner8_kernel(const uint8_t *__restrict__ bases, uint64_t n_bases, const uint32_t *__restrict__ flags,
                          int k, SkmStage st, Counters *__restrict__ ctr, unsigned long long *__restrict__ kmer_count) {
    __shared__ uint32_t s_words[EX_THREADS + 2];
    __shared__ uint32_t s_flags[EX_THREADS / 2 + 2];
    __shared__ unsigned long long s_wrec[EX_THREADS / 32][2];    // per-warp record totals (packed)
    __shared__ unsigned long long s_wkm[EX_THREADS / 32][2];     // per-warp k-mer totals (packed)
    __shared__ uint32_t s_base[8];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t G = st.n_regions;
    const uint64_t n_tiles = (((n_bases + 15) >> 4) + EX_THREADS - 1) / EX_THREADS;
    uint32_t bad = 0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TileWord t = load_tile_word<EX_THREADS>(bases, n_bases, flags, tile, k, s_words, s_flags, bad);
        const uint32_t w0 = t.w0, w1 = t.w1, w2 = t.w2;
        uint32_t valid = 0;
        uint32_t mh[16];
        if (t.active) {
            const uint64_t span = (k > 1) ? ((1ULL << (k - 1)) - 1ULL) : 1ULL;
#pragma unroll
            for (int j = 0; j < 16; j++) valid |= ((((t.fbits >> j) & span) == 0 && (long long)j <= t.limit) ? 1u : 0u) << j;
        }
        if (valid) minhash_of_word(w0, w1, w2, k, minimizer_len(k), mh);
        // runs of consecutive valid k-mers with the same minimizer hash; f(j0, len, minhash)
        auto for_each_run = [&](auto f) {
            uint32_t run_start = 0, run_mh = 0;
            bool in_run = false;
#pragma unroll
            for (int j = 0; j <= 16; j++) {
                const bool v = j < 16 && ((valid >> j) & 1);
                const uint32_t mhj = mh[j < 16 ? j : 15];
                if (in_run && (!v || mhj != run_mh)) { f(run_start, (uint32_t)j - run_start, run_mh); in_run = false; }
                if (v && !in_run) { in_run = true; run_start = (uint32_t)j; run_mh = mhj; }
            }
        };
        // pass 1: records and k-mers per owner
        unsigned long long rc_lo = 0, rc_hi = 0, km_lo = 0, km_hi = 0;
        if (valid) for_each_run([&](uint32_t, uint32_t len, uint32_t mhv) {
            const uint32_t o = owner_of_minhash(mhv, G);
            pk_add(rc_lo, rc_hi, o, 1u); pk_add(km_lo, km_hi, o, len);
        });
        unsigned long long in_lo = rc_lo, in_hi = rc_hi;               // inclusive warp scan of the record counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long a = __shfl_up_sync(0xffffffffu, in_lo, o), b = __shfl_up_sync(0xffffffffu, in_hi, o);
            if (lane >= (uint32_t)o) { in_lo += a; in_hi += b; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { km_lo += __shfl_xor_sync(0xffffffffu, km_lo, o); km_hi += __shfl_xor_sync(0xffffffffu, km_hi, o); }
        if (lane == 31) { s_wrec[warp][0] = in_lo; s_wrec[warp][1] = in_hi; }
        if (lane == 0) { s_wkm[warp][0] = km_lo; s_wkm[warp][1] = km_hi; }
        __syncthreads();
        unsigned long long wb_lo = 0, wb_hi = 0, tot_lo = 0, tot_hi = 0, kt_lo = 0, kt_hi = 0;
#pragma unroll
        for (int wv = 0; wv < EX_THREADS / 32; wv++) {
            if (wv < (int)warp) { wb_lo += s_wrec[wv][0]; wb_hi += s_wrec[wv][1]; }
            tot_lo += s_wrec[wv][0]; tot_hi += s_wrec[wv][1];
            kt_lo += s_wkm[wv][0]; kt_hi += s_wkm[wv][1];
        }
        if (tid < G) {                                                  // one reservation per (tile, owner)
            const uint32_t n = pk_get(tot_lo, tot_hi, tid);
            s_base[tid] = n ? atomicAdd(&st.cursor[tid], n) : 0u;
            const uint32_t km = pk_get(kt_lo, kt_hi, tid);
            if (km) atomicAdd(&kmer_count[tid], (unsigned long long)km);
        }
        __syncthreads();
        // pass 2: write the records
        if (valid) {
            const unsigned long long ex_lo = wb_lo + in_lo - rc_lo, ex_hi = wb_hi + in_hi - rc_hi;   // exclusive prefix of this thread
            unsigned long long run_lo = 0, run_hi = 0;
            for_each_run([&](uint32_t j0, uint32_t len, uint32_t mhv) {
                const uint32_t o = owner_of_minhash(mhv, G);
                const uint64_t pos = (uint64_t)s_base[o] + pk_get(ex_lo, ex_hi, o) + pk_get(run_lo, run_hi, o);
                pk_add(run_lo, run_hi, o, 1u);
                if (pos < st.seg_cap) {
                    const int sh = 2 * (int)j0;
                    uint4 rec;
                    rec.x = sh ? __funnelshift_l(w1, w0, sh) : w0;
                    rec.y = sh ? __funnelshift_l(w2, w1, sh) : w1;
                    rec.z = ((sh ? (w2 << sh) : w2) & ~15u) | (len - 1);
                    rec.w = mhv;
                    st.recs[(uint64_t)o * st.seg_cap + pos] = rec;
                }
            });
        }
        __syncthreads();
    }
    if (__any_sync(0xffffffffu, bad != 0) && lane_id() == 0) atomicAdd(&ctr->bad_chars, 1ULL);
}

// Receive side of the shard exchange: records that arrived from other GPUs are filed under the
// table region of their minimizer (streaming: 16 B in, one cursor atomic, 16 B out).
template <class Tab>
__global__ void __launch_bounds__(256)
skm_restage_kernel(const uint4 *__restrict__ in, uint64_t n, int k, SkmStage st, Tab tb, Counters *__restrict__ ctr) {
    uint32_t claimed = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 rec = ld_nc_u128(&in[i]);
        const uint32_t region = region_of_minhash(rec.w, st.n_regions);
        const uint32_t pos = atomicAdd(&st.cursor[region], 1u);
        if (pos < st.seg_cap) st.recs[(uint64_t)region * st.seg_cap + pos] = rec;
        else claimed += skm_count_direct(rec, region, st.region_shift, st.win, k, tb);
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

// Phase B for super-k-mer records, L2 flavour: blocks_per_region consecutive CTAs own one region (CTAs
// start in index order, so only a few regions are live in L2 at a time); the k-mer instances of 32 records
// are dealt out evenly over the lanes of a warp (records hold 1..16 k-mers, 6 on average) and every lane
// upserts one instance per step with global atomics.  `put(home_slot, key)` is the table operation.
template <class Put>
__device__ __forceinline__ void skm_expand_records(const uint4 *__restrict__ recs, uint64_t n, uint64_t first, uint64_t stride,
                                                   int k, uint4 (*s_rec)[32], uint32_t (*s_pre)[33], Put put) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rs = 64 - 2 * k;
    uint64_t base = first + (uint64_t)warp * 32;
    uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
    if (base + lane < n) nxt = ld_nc_u128(&recs[base + lane]);         // software pipeline: the next 32 records are in
    for (; base < n; base += stride) {                                   // flight while these are expanded (the records
        const uint64_t i = base + lane;                                  // may live in a peer GPU's memory)
        const uint4 r = nxt;
        const uint32_t len = i < n ? (r.z & 15u) + 1u : 0u;
        if (i + stride < n) nxt = ld_nc_u128(&recs[i + stride]);
        uint32_t incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
        s_rec[warp][lane] = r;
        s_pre[warp][lane + 1] = incl;
        if (lane == 0) s_pre[warp][0] = 0;
        __syncwarp();
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        for (uint32_t t = lane; t < total; t += 32) {
            uint32_t lo = 0, hi = 31;                       // largest record index with prefix <= t
#pragma unroll
            for (int step = 0; step < 5; step++) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (s_pre[warp][mid] <= t) lo = mid; else hi = mid - 1;
            }
            const uint32_t off = t - s_pre[warp][lo];
            const uint4 q = s_rec[warp][lo];
            const uint32_t w2 = q.z & ~15u;
            const uint32_t h32 = __funnelshift_l(q.y, q.x, 2 * off);
            const uint32_t l32 = __funnelshift_l(w2, q.y, 2 * off);
            const uint64_t fw = (((uint64_t)h32 << 32) | l32) >> rs;
            const uint64_t rc = revcomp64(fw, k);
            put(fw < rc ? fw : rc, q.w);
        }
        __syncwarp();
    }
}

template <class Tab>
__global__ void __launch_bounds__(256)
drain_skm_kernel(SkmStage st, uint32_t blocks_per_region, int k, Tab tb, Counters *__restrict__ ctr) {
    __shared__ uint4 s_rec[8][32];
    __shared__ uint32_t s_pre[8][33];
    const uint32_t region = blockIdx.x / blocks_per_region;
    const uint32_t sub = blockIdx.x % blocks_per_region;
    uint64_t n = st.cursor[region];
    if (n > st.seg_cap) n = st.seg_cap;
    const uint64_t region_base = (uint64_t)region << st.region_shift;
    const uint64_t slot_mask = (1ull << st.region_shift) - 1ull;
    uint32_t claimed = 0;
    skm_expand_records(st.recs + (uint64_t)region * st.seg_cap, n, (uint64_t)sub * 256, (uint64_t)blocks_per_region * 256, k, s_rec, s_pre,
                       [&](uint64_t key, uint32_t) { claimed += tb.upsert1_at(region_base | (mix64(key) & slot_mask), key, st.region_shift, st.win) ? 1u : 0u; });
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if ((threadIdx.x & 31) == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

// ------------------------------------------------------------------------------------------
// Phase B, shared-memory flavour (default for k <= 31): regions are small enough (<= 2^13 slots,
// 2^12 = 64 KiB by default) for ONE CTA to hold the whole region of the table in shared memory.
// The CTA streams the region in (coalesced 16-byte loads), expands the region's super-k-mer records
// and upserts them with shared-memory atomics (no L2 round trip per k-mer), then streams the region
// back.  HBM sees pure streaming traffic: the table once in and once out per drain, the records once.
//
// Placement is the windowed rule of placed_upsert_at: a k-mer whose primary window (SMEM_WIN slots,
// wrapping inside the region) is full of other keys belongs at its secondary position somewhere else
// in the table, which this CTA cannot touch.  Such instances are SPILLED: up to SMEM_SPILL_MAX keys
// per CTA go to a global list; a CTA with more (an overfull region: minimizer skew) commits what it
// counted and reports the region as DIRTY instead.  drain_fallback_kernel, launched right behind on
// the same stream, upserts the listed keys at their secondary positions and re-expands the records
// of dirty regions: an instance whose key now sits in its primary window was counted here (a window
// never gets free slots back, so either all instances of a key were counted in shared memory or
// none), every other instance goes to the secondary position.  Skew costs speed, never results.
// ------------------------------------------------------------------------------------------
constexpr int SMEM_MAX_SHIFT = 13;
constexpr uint32_t SMEM_SPILL_MAX = 256;       // spilled instances a CTA may list
constexpr uint32_t SMEM_WIN = 64;              // primary window of the windowed placement

struct SpillBuf {
    unsigned long long *keys;   // spilled instances (their secondary position derives from the key)
    unsigned int *cursor;       // entries used
    unsigned int *n_dirty;
    uint32_t *dirty;            // region ids, n_regions entries
    uint32_t cap;
};

__global__ void __launch_bounds__(256, 3)
drain_smem_kernel(SkmStage st, int k, Slot *__restrict__ tab, SpillBuf sp, Counters *__restrict__ ctr) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint4 *s_tab = reinterpret_cast<uint4 *>(smem_raw);                       // the region: x,y = key, z = count
    const uint32_t S = 1u << st.region_shift;
    unsigned long long *s_spill = reinterpret_cast<unsigned long long *>(s_tab + S);
    __shared__ uint4 s_rec[8][32];
    __shared__ uint32_t s_pre[8][33];
    __shared__ uint32_t s_nspill, s_claimed, s_base;
    __shared__ int s_mode;                            // 0: no spills, 1: spills listed, 2: region dirty
    const uint32_t region = blockIdx.x, tid = threadIdx.x;
    uint64_t n = st.cursor[region];
    if (n == 0) return;
    if (n > st.seg_cap) n = st.seg_cap;
    const uint64_t region_base = (uint64_t)region << st.region_shift;
    uint4 *__restrict__ g_tab = reinterpret_cast<uint4 *>(tab + region_base);
    for (uint32_t i = tid; i < S; i += 256) s_tab[i] = ld_nc_u128(&g_tab[i]);
    if (tid == 0) { s_nspill = 0; s_claimed = 0; }
    __syncthreads();

    uint32_t claimed = 0;
    const uint32_t slot_mask = S - 1u;
    const uint32_t win = SMEM_WIN < S ? SMEM_WIN : S;
    skm_expand_records(st.recs + (uint64_t)region * st.seg_cap, n, 0, 256, k, s_rec, s_pre, [&](uint64_t key, uint32_t) {
        const uint32_t klo = (uint32_t)key, khi = (uint32_t)(key >> 32);
        uint32_t i = (uint32_t)mix64(key) & slot_mask;
        for (uint32_t step = 0; step < win; step++, i = (i + 1) & slot_mask) {
            const uint4 v = s_tab[i];
            if (v.x == klo && v.y == khi) { if (v.z < MAX_COUNT) atomicAdd(&s_tab[i].z, 1u); return; }
            if ((v.x & v.y) == 0xFFFFFFFFu) {
                const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long *>(&s_tab[i]), EMPTY_KEY, (unsigned long long)key);
                if (prev == EMPTY_KEY) { atomicAdd(&s_tab[i].z, 1u); claimed++; return; }
                if (prev == key) { atomicAdd(&s_tab[i].z, 1u); return; }
            }
        }
        const uint32_t at = atomicAdd(&s_nspill, 1u);
        if (at < SMEM_SPILL_MAX) s_spill[at] = key;
    });
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if ((tid & 31) == 0 && claimed) atomicAdd(&s_claimed, claimed);
    __syncthreads();

    if (tid == 0) {
        const uint32_t ns = s_nspill;
        int mode = ns ? 2 : 0;
        uint32_t base = 0;
        if (ns && ns <= SMEM_SPILL_MAX) {             // reserve room in the global list (all or nothing)
            unsigned int old = *(volatile unsigned int *)sp.cursor;
            for (;;) {
                if (old + ns > sp.cap) break;
                const unsigned int seen = atomicCAS(sp.cursor, old, old + ns);
                if (seen == old) { base = old; mode = 1; break; }
                old = seen;
            }
        }
        if (mode == 2) sp.dirty[atomicAdd(sp.n_dirty, 1u)] = region;
        else st.cursor[region] = 0;
        if (s_claimed) atomicAdd(&ctr->distinct, (unsigned long long)s_claimed);
        s_mode = mode; s_base = base;
    }
    __syncthreads();
    if (s_mode == 1) { const uint32_t ns = s_nspill; for (uint32_t i = tid; i < ns; i += 256) sp.keys[s_base + i] = s_spill[i]; }
    for (uint32_t i = tid; i < S; i += 256) {
        uint4 v = s_tab[i];
        if (v.z > MAX_COUNT) v.z = MAX_COUNT;
        g_tab[i] = v;
    }
}

// Runs right after drain_smem_kernel on the same stream (every region is back in HBM): listed spills
// go to their secondary positions; the records of dirty regions are expanded again and the instances
// that were NOT counted in shared memory (key absent from its primary window) follow them.
__global__ void __launch_bounds__(256)
drain_fallback_kernel(SkmStage st, int k, Slot *__restrict__ tab, uint64_t cap, SpillBuf sp, Counters *__restrict__ ctr) {
    __shared__ uint4 s_rec[8][32];
    __shared__ uint32_t s_pre[8][33];
    uint32_t claimed = 0;
    uint32_t ns = *sp.cursor;
    if (ns > sp.cap) ns = sp.cap;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < ns; i += gridDim.x * 256) {
        const unsigned long long key = sp.keys[i];
        claimed += secondary_upsert<true>(tab, cap, SMEM_WIN, key, 1u) ? 1u : 0u;
    }
    const uint32_t nd = *sp.n_dirty;
    const uint64_t slot_mask = (1ull << st.region_shift) - 1ull;
    const uint32_t win = (uint64_t)SMEM_WIN < slot_mask + 1 ? SMEM_WIN : (uint32_t)(slot_mask + 1);
    for (uint32_t f = blockIdx.x; f < nd; f += gridDim.x) {
        const uint32_t region = sp.dirty[f];
        uint64_t n = st.cursor[region];
        if (n > st.seg_cap) n = st.seg_cap;
        const uint64_t region_base = (uint64_t)region << st.region_shift;
        skm_expand_records(st.recs + (uint64_t)region * st.seg_cap, n, 0, 256, k, s_rec, s_pre, [&](uint64_t key, uint32_t) {
            uint64_t off = mix64(key) & slot_mask;
            for (uint32_t step = 0; step < win; step++, off = (off + 1) & slot_mask) {
                const unsigned long long cur = ld_cg_u64x2(&tab[region_base | off]).x;
                if (cur == key) return;                               // counted in shared memory
                if (cur == EMPTY_KEY) {                               // cannot happen after a drain; stay exact anyway
                    claimed += placed_upsert_at<true>(tab, cap, st.region_shift, SMEM_WIN, region_base | off, key, 1u) ? 1u : 0u;
                    return;
                }
            }
            claimed += secondary_upsert<true>(tab, cap, SMEM_WIN, key, 1u) ? 1u : 0u;
        });
        __syncthreads();
        if (threadIdx.x == 0) st.cursor[region] = 0;
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if ((threadIdx.x & 31) == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

// ------------------------------------------------------------------------------------------
// Peer-memory shard exchange (multi-GPU): compute + collective in ONE kernel.
// Every GPU stages its super-k-mer records in its OWN memory, bucketed by (owner shard, coarse
// bucket of the minimizer hash) -- extract_skm_kernel<2>.  The owner's drain then reads "its"
// segments straight out of every peer's staging buffer with P2P loads over NVLink/NVSwitch
// (pointers opened with CUDA IPC) and upserts them into its table: no send/receive buffers, no
// NCCL data path, no re-staging pass, and the transfer overlaps the upserts record by record
// (skm_expand_records keeps the next 32 records of the warp in flight).
// The coarse bucket is the top bits of the same hash the table region derives from, so the records
// of bucket b fall into a contiguous 1/B of the owner's table whatever its size: consecutive CTAs
// work on the same bucket and its table slice stays L2-resident, as in drain_skm_kernel.
// ------------------------------------------------------------------------------------------
constexpr int P2P_MAX_PEERS = 16;
struct P2PPeers {
    const uint4 *recs[P2P_MAX_PEERS];             // peer s: segment (owner << log2_buckets | bucket) of seg_cap records
    const unsigned int *cursor[P2P_MAX_PEERS];
    uint64_t seg_cap;
    uint32_t n_peers, me;
    int log2_buckets;
};

template <class Tab>
__global__ void __launch_bounds__(256)
drain_p2p_kernel(P2PPeers pp, uint32_t bucket0, uint32_t blocks_per_bucket, int k, Tab tb, uint32_t n_regions, int region_shift, uint32_t win,
                 Counters *__restrict__ ctr) {
    __shared__ uint4 s_rec[8][32];
    __shared__ uint32_t s_pre[8][33];
    const uint32_t bucket = bucket0 + blockIdx.x / blocks_per_bucket;
    const uint32_t sub = blockIdx.x % blocks_per_bucket;
    const uint64_t seg = ((uint64_t)pp.me << pp.log2_buckets) | bucket;
    const uint64_t slot_mask = (1ull << region_shift) - 1ull;
    uint32_t claimed = 0;
    for (uint32_t j = 0; j < pp.n_peers; j++) {
        const uint32_t s = (pp.me + j) % pp.n_peers;            // start at home, then round the peers: spreads the NVLink load
        uint64_t n = pp.cursor[s][seg];
        if (n > pp.seg_cap) n = pp.seg_cap;
        skm_expand_records(pp.recs[s] + seg * pp.seg_cap, n, (uint64_t)sub * 256, (uint64_t)blocks_per_bucket * 256, k, s_rec, s_pre,
                           [&](uint64_t key, uint32_t mh) {
                               const uint64_t region = region_of_minhash(mh, n_regions);
                               claimed += tb.upsert1_at((region << region_shift) | (mix64(key) & slot_mask), key, region_shift, win) ? 1u : 0u;
                           });
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if ((threadIdx.x & 31) == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

// receive side of the shard exchange / generic "count these keys" (direct random upserts)
__global__ void __launch_bounds__(256)
count_keys_kernel(const unsigned long long *__restrict__ keys, uint64_t n, Slot *__restrict__ tab, TableGeom g,
                  Counters *__restrict__ ctr) {
    uint32_t claimed = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        claimed += placed_upsert_at<true>(tab, g.cap, g.region_shift, g.minimizer ? g.win : 0u, geom_home(g, key), key, 1u) ? 1u : 0u;
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

// SoA (minimizer-placed) flavours of the generic table kernels
__global__ void __launch_bounds__(256)
count_keys_soa_kernel(const unsigned long long *__restrict__ keys, uint64_t n, TabSoA tb, TableGeom g, Counters *__restrict__ ctr) {
    uint32_t claimed = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        claimed += soa_upsert1_at(tb, geom_home(g, key), key) ? 1u : 0u;
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}
__global__ void __launch_bounds__(256)
rehash_soa_kernel(TabSoA old_t, TabSoA new_t, TableGeom g) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < old_t.cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = old_t.keys[i];
        if (key != EMPTY_KEY) {
            const uint32_t c = old_t.counts[i];
            soa_upsert_n_at(new_t, geom_home(g, key), key, c < MAX_COUNT ? c : MAX_COUNT);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4: histogram over ALL entries (src/io/IOUtils.java:59), filter count > b, compaction
// ------------------------------------------------------------------------------------------
constexpr int HIST_SMEM_BINS = 2048;

// one slot as (key lo, key hi, count) whatever the layout
template <bool SOA>
__device__ __forceinline__ uint4 load_slot(const Slot *__restrict__ tab, const TabSoA &tb, uint64_t i) {
    if (SOA) {
        const unsigned long long k = tb.keys[i];
        return make_uint4((uint32_t)k, (uint32_t)(k >> 32), tb.counts[i], 0u);
    }
    return ld_nc_u128(&tab[i]);
}

template <bool SOA>
__global__ void __launch_bounds__(256)
table_hist_kernel(const Slot *__restrict__ tab, TabSoA tb, uint64_t cap, unsigned long long *__restrict__ hist) {
    __shared__ uint32_t s_hist[HIST_SMEM_BINS];
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 s = load_slot<SOA>(tab, tb, i);
        if ((s.x & s.y) != 0xFFFFFFFFu) {
            const uint32_t c = s.z < MAX_COUNT ? s.z : MAX_COUNT;
            if (c < HIST_SMEM_BINS) atomicAdd(&s_hist[c], 1u);
            else atomicAdd(&hist[c], 1ULL);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

__global__ void __launch_bounds__(256)
table_compact_kernel(const Slot *__restrict__ tab, uint64_t cap, uint32_t threshold,
                     unsigned long long *__restrict__ out_keys, uint16_t *__restrict__ out_counts,
                     uint64_t out_cap, Counters *__restrict__ ctr) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_iter = (cap + stride - 1) / stride;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < n_iter; it++, i += stride) {
        bool good = false;
        unsigned long long key = 0; uint32_t c = 0;
        if (i < cap) {
            const uint4 s = ld_nc_u128(&tab[i]);
            key = ((unsigned long long)s.y << 32) | s.x;
            c = s.z < MAX_COUNT ? s.z : MAX_COUNT;
            good = key != EMPTY_KEY && c > threshold;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, good);
        if (!m) continue;
        unsigned long long base = 0;
        if (lane_id() == 0) base = atomicAdd(&ctr->n_good, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (good) {
            const uint64_t at = base + __popc(m & lanemask_lt());
            if (at < out_cap) { out_keys[at] = key; out_counts[at] = (uint16_t)c; }
        }
    }
}

// One pass over the table: histogram of ALL entries + compaction of the entries with
// count > threshold.  Output positions are reserved with ONE global atomic per block iteration
// (warp ballots -> shared prefix), so the cursor is not a serialisation point.
template <bool SOA>
__global__ void __launch_bounds__(256)
table_scan_kernel(const Slot *__restrict__ tab, TabSoA tb, uint64_t cap, uint32_t threshold, unsigned long long *__restrict__ hist,
                  unsigned long long *__restrict__ out_keys, uint16_t *__restrict__ out_counts, uint64_t out_cap,
                  Counters *__restrict__ ctr) {
    constexpr int U = 8;                                   // slots per thread per round: 8 x 16 B in flight
    __shared__ uint32_t s_hist[HIST_SMEM_BINS];
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_base;
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t round_slots = (uint64_t)256 * U;
    for (uint64_t base = (uint64_t)blockIdx.x * round_slots; base < cap; base += (uint64_t)gridDim.x * round_slots) {
        uint4 s[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t i = base + (uint64_t)u * 256 + threadIdx.x;
            s[u] = i < cap ? load_slot<SOA>(tab, tb, i) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
        }
        uint32_t good_mask = 0;
#pragma unroll
        for (int u = 0; u < U; u++) {
            if ((s[u].x & s[u].y) != 0xFFFFFFFFu) {
                const uint32_t c = s[u].z < MAX_COUNT ? s[u].z : MAX_COUNT;
                s[u].z = c;
                if (c < HIST_SMEM_BINS) atomicAdd(&s_hist[c], 1u); else atomicAdd(&hist[c], 1ULL);
                if (c > threshold) good_mask |= 1u << u;
            }
        }
        const uint32_t mine = __popc(good_mask);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int wv = 0; wv < 8; wv++) { const uint32_t cc = s_warp[wv]; if (wv < (int)warp) before += cc; total += cc; }
        if (threadIdx.x == 0 && total) s_base = atomicAdd(&ctr->n_good, (unsigned long long)total);
        __syncthreads();
        if (mine) {
            uint64_t at = s_base + before + incl - mine;
#pragma unroll
            for (int u = 0; u < U; u++) {
                if ((good_mask >> u) & 1) {
                    if (at < out_cap) {
                        out_keys[at] = ((unsigned long long)s[u].y << 32) | s[u].x;
                        out_counts[at] = (uint16_t)s[u].z;
                    }
                    at++;
                }
            }
        }
    }
    __syncthreads();
    for (int i2 = threadIdx.x; i2 < HIST_SMEM_BINS; i2 += blockDim.x)
        if (s_hist[i2]) atomicAdd(&hist[i2], (unsigned long long)s_hist[i2]);
}

// sorted (key,count) -> big-endian 10-byte records (src/io/IOUtils.java:61-65: writeLong, writeShort)
__global__ void __launch_bounds__(256)
records_kernel(const unsigned long long *__restrict__ keys, const uint16_t *__restrict__ counts, uint64_t n,
               uint16_t *__restrict__ out /* 5 x u16 per record */) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        const uint32_t c = counts[i];
        uint16_t *o = out + 5 * i;
        // big-endian bytes, emitted as little-endian u16 stores of byte-swapped halves
        const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
        o[0] = (uint16_t)__byte_perm(hi, 0, 0x0023);   // bytes 3,2 of hi -> [b3][b2] in memory order
        o[1] = (uint16_t)__byte_perm(hi, 0, 0x0001);
        o[2] = (uint16_t)__byte_perm(lo, 0, 0x0023);
        o[3] = (uint16_t)__byte_perm(lo, 0, 0x0001);
        o[4] = (uint16_t)__byte_perm(c, 0, 0x0001);
    }
}

// sort variant: run-length encode a sorted key array.  heads[i] = 1 iff keys[i] starts a run.
__global__ void __launch_bounds__(256)
rle_mark_kernel(const unsigned long long *__restrict__ keys, uint64_t n, unsigned long long *__restrict__ n_runs_block) {
    // counts run heads per block (block b covers a contiguous slice) for the exclusive scan
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint64_t per_block = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = per_block * blockIdx.x;
    const uint64_t hi = lo + per_block < n ? lo + per_block : n;
    uint32_t local = 0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
        local += (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane_id() == 0 && local) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) n_runs_block[blockIdx.x] = s_cnt;
}

// After an exclusive scan of n_runs_block (-> run_base[b]): each block walks its slice in order and
// writes (key, run length) for every run head; run length = distance to the next head, found by
// scanning forward (runs are short on average; heavy runs are walked by one thread but only once).
__global__ void __launch_bounds__(256)
rle_write_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ weights /* may be NULL */,
                 uint64_t n, const unsigned long long *__restrict__ run_base,
                 unsigned long long *__restrict__ out_keys, uint32_t *__restrict__ out_counts) {
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
            unsigned long long len = 0;
            for (uint64_t j = i; j < n && keys[j] == key; j++) {
                len += weights ? weights[j] : 1u;
                if (len >= MAX_COUNT && !weights) { /* saturated: skip ahead cheaply */ }
            }
            out_keys[at] = key;
            out_counts[at] = len > MAX_COUNT ? MAX_COUNT : (uint32_t)len;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
}

// histogram + filter over an already RLE'd sorted (key,count) array (sort variant emit)
__global__ void __launch_bounds__(256)
pairs_hist_kernel(const uint32_t *__restrict__ counts, uint64_t n, unsigned long long *__restrict__ hist) {
    __shared__ uint32_t s_hist[HIST_SMEM_BINS];
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = counts[i] < MAX_COUNT ? counts[i] : MAX_COUNT;
        if (c < HIST_SMEM_BINS) atomicAdd(&s_hist[c], 1u); else atomicAdd(&hist[c], 1ULL);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HIST_SMEM_BINS; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

// ------------------------------------------------------------------------------------------
// K6/K7: features-calculator
// ------------------------------------------------------------------------------------------
// hm.put(kmer, 0) for every component k-mer (FeaturesCalculatorMain.java:97-103)
__global__ void __launch_bounds__(256)
fc_build_kernel(const unsigned long long *__restrict__ keys, uint64_t n, FcSlot *__restrict__ tab, uint64_t cap,
                Counters *__restrict__ ctr, uint32_t *__restrict__ bloom, uint32_t bmask) {
    uint32_t claimed = 0;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[t];
        if (bmask) { const uint32_t b = (uint32_t)mix64(key) & bmask; atomicOr(&bloom[b >> 5], 1u << (b & 31u)); }
        uint64_t i = home_slot(key, cap);
        for (;;) {
            unsigned long long cur = ld_cg_u64x2(&tab[i]).x;
            if (cur == EMPTY_KEY) {
                cur = atomicCAS(&tab[i].key, EMPTY_KEY, key);
                if (cur == EMPTY_KEY) { claimed++; break; }
            }
            if (cur == key) break;
            if (++i == cap) i = 0;
        }
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

__global__ void __launch_bounds__(256)
fc_clear_kernel(FcSlot *__restrict__ tab, uint64_t cap, int keys_too) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        if (keys_too) tab[i].key = EMPTY_KEY;
        tab[i].acc = 0;
    }
}

__device__ __forceinline__ void load_record(const uint8_t *__restrict__ rec, unsigned long long &key, int &freq) {
    // 10-byte big-endian record at a 2-byte aligned address (KmersLoadWorker.java:24-27)
    const uint16_t *p = reinterpret_cast<const uint16_t *>(rec);
    unsigned long long kk = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { const uint32_t h = p[j]; kk = (kk << 16) | (uint32_t)__byte_perm(h, 0, 0x4401); }
    key = kk;
    const uint32_t h = p[4];
    freq = (int)(short)__byte_perm(h, 0, 0x4401);
}

// KmersPresenceWorker.processKmer (src/io/IOUtils.java:583-587)
__global__ void __launch_bounds__(256)
fc_records_kernel(const uint8_t *__restrict__ recs, uint64_t n, FcSlot *__restrict__ tab, uint64_t cap,
                  const uint32_t *__restrict__ bloom, uint32_t bmask) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long key; int freq;
        load_record(recs + 10 * t, key, freq);
        fc_accumulate(tab, cap, key, (long long)freq, bloom, bmask);
    }
}

// the same accumulate for (key, count) pairs that are still on the device (the counter's emit arrays): the
// kmer-counter-many -> features-calculator hand-over without the round trip through a .kmers.bin file
__global__ void __launch_bounds__(256)
fc_pairs_kernel(const unsigned long long *__restrict__ keys, const uint16_t *__restrict__ counts, uint64_t n,
                FcSlot *__restrict__ tab, uint64_t cap, const uint32_t *__restrict__ bloom, uint32_t bmask) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x)
        fc_accumulate(tab, cap, keys[t], (long long)(short)counts[t], bloom, bmask);
}

// Kmers2HMWorker.processKmer with threshold 0 (src/io/IOUtils.java:249-257): selected[key] =
// sat_add16(selected[key], freq) for freq > 0.  The selected set reuses Slot.
__global__ void __launch_bounds__(256)
fc_selected_kernel(const uint8_t *__restrict__ recs, uint64_t n, Slot *__restrict__ tab, uint64_t cap,
                   Counters *__restrict__ ctr) {
    uint32_t claimed = 0;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long key; int freq;
        load_record(recs + 10 * t, key, freq);
        if (freq > 0) claimed += table_upsert_n(tab, cap, key, (uint32_t)freq) ? 1u : 0u;
    }
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if (lane_id() == 0 && claimed) atomicAdd(&ctr->distinct, (unsigned long long)claimed);
}

__device__ __forceinline__ long long fc_lookup(const FcSlot *__restrict__ tab, uint64_t cap, unsigned long long key) {
    uint64_t i = home_slot(key, cap);
    for (;;) {
        const ulonglong2 s = ld_cg_u64x2(&tab[i]);
        if (s.x == key) return (long long)s.y;
        if (s.x == EMPTY_KEY) return 0;                     // getWithZero
        if (++i == cap) i = 0;
    }
}
__device__ __forceinline__ uint32_t sel_lookup(const Slot *__restrict__ tab, uint64_t cap, unsigned long long key) {
    uint64_t i = home_slot(key, cap);
    for (;;) {
        const ulonglong2 s = ld_cg_u64x2(&tab[i]);
        if (s.x == key) return (uint32_t)s.y;
        if (s.x == EMPTY_KEY) return 0;
        if (++i == cap) i = 0;
    }
}

// buildAndPrintVector inner loop (src/tools/FeaturesCalculatorMain.java:186-204): one warp per
// component, lanes stride over its keys, int64 sums (wrapping like Java long).
__global__ void __launch_bounds__(256)
fc_features_kernel(const unsigned long long *__restrict__ comp_keys, const uint64_t *__restrict__ comp_off,
                   uint32_t n_comp, const FcSlot *__restrict__ tab, uint64_t cap,
                   const Slot *__restrict__ sel, uint64_t sel_cap, long long threshold,
                   long long *__restrict__ vec, unsigned long long *__restrict__ found, unsigned long long *__restrict__ cnt) {
    const uint32_t warps_per_block = blockDim.x >> 5;
    for (uint32_t c = blockIdx.x * warps_per_block + (threadIdx.x >> 5); c < n_comp; c += gridDim.x * warps_per_block) {
        unsigned long long sum = 0, f = 0, n = 0;
        for (uint64_t i = comp_off[c] + lane_id(); i < comp_off[c + 1]; i += 32) {
            const unsigned long long key = comp_keys[i];
            if (sel == nullptr || sel_lookup(sel, sel_cap, key) > 0) {
                const long long v = fc_lookup(tab, cap, key);
                if (v > threshold) { sum += (unsigned long long)v; f++; }
                n++;
            }
        }
        for (int o = 16; o; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            f += __shfl_xor_sync(0xffffffffu, f, o);
            n += __shfl_xor_sync(0xffffffffu, n, o);
        }
        if (lane_id() == 0) { vec[c] = (long long)sum; found[c] = f; cnt[c] = n; }
    }
}

// ------------------------------------------------------------------------------------------
// GUPS-style microbenchmark: random 32-byte-sector accesses over a big table.
//   mode 0: one red.add per update (fire and forget)
//   mode 1: dependent 128-bit load + red.add -- the shape of a table upsert
//   mode 2: 128-bit load only (random sector reads)
//   mode 3: like mode 1, but the table is swept window by window (window_sectors each): block b
//           only touches window b / blocks_per_window, so the live working set is a few windows
//           and stays L2-resident -- the access pattern of the region-blocked upsert.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gups_kernel(unsigned long long *__restrict__ tab, uint64_t n_sectors, uint64_t n_updates, uint64_t seed, int mode,
            uint64_t window_sectors, uint32_t blocks_per_window) {
    unsigned long long sink = 0;
    if (mode == 4 || mode == 5) {
        // split layout experiment: keys (8 B) in the first half of the buffer, counts (4 B) in the second;
        // windowed like mode 3.  mode 4: ld key + red count (different sectors); mode 5: ld key only + CAS key
        const uint64_t n_slots = n_sectors * 2;                   // 16 B of buffer per slot: 8 B key + 4 B count (+4 unused)
        unsigned long long *keys = tab;
        unsigned int *counts = reinterpret_cast<unsigned int *>(tab + n_slots);
        const uint64_t win_slots = window_sectors * 2;
        const uint64_t n_windows = (n_slots + win_slots - 1) / win_slots;
        const uint64_t per_block = n_updates / ((uint64_t)n_windows * blocks_per_window) + 1;
        for (uint64_t win = blockIdx.x / blocks_per_window; win < n_windows; win += gridDim.x / blocks_per_window) {
            const uint64_t w0 = win * win_slots;
            const uint64_t wn = w0 + win_slots <= n_slots ? win_slots : n_slots - w0;
            const uint64_t salt = (win * blocks_per_window + blockIdx.x % blocks_per_window) * per_block;
            for (uint64_t i = threadIdx.x; i < per_block; i += blockDim.x) {
                const uint64_t sl = w0 + mulhi64(mix64(salt + i + seed), wn);
                unsigned long long v;
                asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(keys + sl));
                if (mode == 4) { if (v != 0x123456789ULL) atomicAdd(&counts[sl], 1u); else sink += v; }
                else { if (v == 0) atomicCAS(&keys[sl], 0ull, sl + 1); else sink += v; }
            }
        }
    } else if (mode == 3) {
        const uint64_t n_windows = (n_sectors + window_sectors - 1) / window_sectors;
        const uint64_t per_block = n_updates / ((uint64_t)n_windows * blocks_per_window) + 1;
        for (uint64_t win = blockIdx.x / blocks_per_window; win < n_windows; win += gridDim.x / blocks_per_window) {
            const uint64_t w0 = win * window_sectors;
            const uint64_t wn = w0 + window_sectors <= n_sectors ? window_sectors : n_sectors - w0;
            const uint64_t salt = (win * blocks_per_window + blockIdx.x % blocks_per_window) * per_block;
            for (uint64_t i = threadIdx.x; i < per_block; i += blockDim.x) {
                const uint64_t sct = w0 + mulhi64(mix64(salt + i + seed), wn);
                unsigned long long *p = tab + 4 * sct;
                const ulonglong2 v = ld_cg_u64x2(p);
                if (v.x != 0x123456789ULL) atomicAdd((unsigned int *)(p + 1), 1u); else sink += v.y;
            }
        }
    } else {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_updates; i += (uint64_t)gridDim.x * blockDim.x) {
            const uint64_t sct = mulhi64(mix64(i + seed), n_sectors);
            unsigned long long *p = tab + 4 * sct;
            if (mode == 1) {
                const ulonglong2 v = ld_cg_u64x2(p);
                if (v.x != 0x123456789ULL) atomicAdd((unsigned int *)(p + 1), 1u); else sink += v.y;
            } else if (mode == 2) {
                const ulonglong2 v = ld_cg_u64x2(p);
                sink += v.x ^ v.y;
            } else {
                atomicAdd((unsigned int *)(p + 1), 1u);
            }
        }
    }
    if (sink == 0xdeadbeefULL) tab[0] = sink;
}

}  // namespace mfkc
