// components.cuh -- component-cutter's graph half on the device: connected components of the k-mer graph with the size
// window [b1, b2] and the re-split of big components at the next frequency threshold.
//
// Reference: ComponentsBuilder.splitStrategy / run / Task.run / findAllComponents / bfs
// (src/algo/ComponentsBuilder.java:24-31,58-84,157-181,198-269) over KmerOperations.possibleNeighbours
// (src/algo/KmerOperations.java:9-26).  The reference runs a breadth-first search from every unvisited map entry, marks
// visited k-mers by negating their value, and queues components larger than b2 for another round among their k-mers of
// value >= threshold + 1 (`nextHM`).  A BFS always walks a component to its end (:244-262), so its result is a
// function of the k-mer SET only -- here it is computed level by level, all components of one frequency threshold at
// once:
//   cc_union_kernel    lock-free union-find (hook the larger root under the smaller with atomicCAS, path halving on
//                      the way up) over the 8 possible neighbours of every active k-mer, found through a hash index
//                      key -> entry number; one pass over the edges gives the components
//   cc_count_kernel    flatten (parent = root) and count the members of every root
//   cc_classify_kernel size < b1: dropped; b1..b2: labelled with (root, threshold) for the output; > b2: the members
//                      of value >= threshold + 1 stay active for the next level
// The map is the key-sorted array pair of a mfkc_kset.  HBM-bound pointer chasing: 8 random index probes (one 16-byte
// slot = half a sector each) per active k-mer and level.
//
// This header uses no warp intrinsics, shared memory or inline PTX on purpose: tests/emu/cc_emu.cpp compiles it for
// the host (one emulated CUDA thread per host thread, real atomics) so that the level logic and the union-find under
// contention are checked against the oracle on CPU-only machines too; the product path is the CUDA build in mfkc.cu
// (mfkc_kset_components_*).
#pragma once
#include "device_common.cuh"

namespace mfkc {

constexpr uint32_t CC_NONE = 0xFFFFFFFFu;

struct CcIndex { const Slot *tab; uint64_t cap; };      // key -> entry number (in Slot::count)

__device__ __forceinline__ int cc_short(uint32_t v) { return (int)(short)(v & 0xFFFFu); }

// KmerOperations.rc (src/algo/KmerOperations.java:63-75): reverse the 2-bit groups, complement, right-align
__device__ __forceinline__ unsigned long long cc_revcomp(unsigned long long x, int k) {
    x = ((x & 0x3333333333333333ULL) << 2) | ((x >> 2) & 0x3333333333333333ULL);
    x = ((x & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL);
    x = ((x & 0x00FF00FF00FF00FFULL) << 8) | ((x >> 8) & 0x00FF00FF00FF00FFULL);
    x = ((x & 0x0000FFFF0000FFFFULL) << 16) | ((x >> 16) & 0x0000FFFF0000FFFFULL);
    x = (x << 32) | (x >> 32);
    return (~x) >> (64 - 2 * k);
}
__device__ __forceinline__ unsigned long long cc_canon(unsigned long long fw, int k) {
    const unsigned long long rc = cc_revcomp(fw, k);
    return fw < rc ? fw : rc;
}

// possibleNeighbours: canonical forms of the 4 right and the 4 left extensions (1 <= k <= 31)
__device__ __forceinline__ unsigned long long cc_neighbour(unsigned long long key, int k, int j) {
    const unsigned long long nuc = (unsigned long long)(j >> 1);
    const unsigned long long mask = (1ULL << (2 * k)) - 1ULL;
    const unsigned long long fw = (j & 1) ? ((key >> 2) | (nuc << (2 * k - 2))) : (((key << 2) | nuc) & mask);
    return cc_canon(fw, k);
}

__global__ void __launch_bounds__(256)
cc_index_build_kernel(const unsigned long long *__restrict__ keys, uint64_t n, Slot *__restrict__ tab, uint64_t cap) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[t];
        uint64_t i = home_slot(key, cap);
        for (;;) {                                            // keys are unique: claim the first free slot
            if (atomicCAS(&tab[i].key, EMPTY_KEY, key) == EMPTY_KEY) { tab[i].count = (uint32_t)t; break; }
            if (++i == cap) i = 0;
        }
    }
}

__device__ __forceinline__ uint32_t cc_lookup(const CcIndex &ix, unsigned long long key) {
    uint64_t i = home_slot(key, ix.cap);
    for (;;) {
        const unsigned long long s = ix.tab[i].key;
        if (s == key) return ix.tab[i].count;
        if (s == EMPTY_KEY) return CC_NONE;
        if (++i == ix.cap) i = 0;
    }
}

// First level: every k-mer with a positive value is a vertex (`startKmer.getValue() > 0`, ComponentsBuilder.java:206)
__global__ void __launch_bounds__(256)
cc_begin_kernel(const uint32_t *__restrict__ vals, uint64_t n, uint8_t *__restrict__ active, uint32_t *__restrict__ label,
                uint32_t *__restrict__ thr_of) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        active[i] = cc_short(vals[i]) > 0 ? 1 : 0;
        label[i] = CC_NONE;
        thr_of[i] = 0;
    }
}

__global__ void __launch_bounds__(256)
cc_level_init_kernel(uint64_t n, uint32_t *__restrict__ parent, uint32_t *__restrict__ size) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        parent[i] = (uint32_t)i;
        size[i] = 0;
    }
}

// Root of x.  Invariant: parent[v] <= v, and only a root (parent[r] == r) is ever re-parented, by a CAS that expects r.
// A stale read therefore yields an ancestor-or-self of x, and the CAS in cc_unite catches a root that is none any more.
__device__ __forceinline__ uint32_t cc_find(uint32_t *parent, uint32_t x) {
    volatile uint32_t *p = parent;
    uint32_t curr = p[x];
    if (curr != x) {
        uint32_t prev = x, next;
        while (curr > (next = p[curr])) {
            p[prev] = next;                                   // path halving: `next` is an ancestor of `prev`
            prev = curr;
            curr = next;
        }
    }
    return curr;
}

// Joins the sets of roots-or-former-roots a and b; returns the root a ends up under (a's side is carried along by the
// caller from neighbour to neighbour)
__device__ __forceinline__ uint32_t cc_unite(uint32_t *parent, uint32_t a, uint32_t b) {
    for (;;) {
        if (a == b) return a;
        if (a < b) {
            const uint32_t old = atomicCAS(&parent[b], b, a);
            if (old == b) return a;
            b = old;                                          // b was no root any more: climb
        } else {
            const uint32_t old = atomicCAS(&parent[a], a, b);
            if (old == a) return b;
            a = old;
        }
    }
}

__global__ void __launch_bounds__(256)
cc_union_kernel(const unsigned long long *__restrict__ keys, uint64_t n, const uint8_t *__restrict__ active, CcIndex ix, int k,
                uint32_t *__restrict__ parent) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        if (!active[t]) continue;
        const unsigned long long key = keys[t];
        uint32_t mine = cc_find(parent, (uint32_t)t);
#pragma unroll 1
        for (int j = 0; j < 8; j++) {
            const uint32_t u = cc_lookup(ix, cc_neighbour(key, k, j));
            if (u == CC_NONE || u == (uint32_t)t || !active[u]) continue;      // absent, itself (poly-A), or not in this level's map
            mine = cc_unite(parent, mine, cc_find(parent, u));
        }
    }
}

__global__ void __launch_bounds__(256)
cc_count_kernel(uint64_t n, const uint8_t *__restrict__ active, uint32_t *__restrict__ parent, uint32_t *__restrict__ size) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        if (!active[t]) continue;
        volatile uint32_t *p = parent;
        uint32_t r = (uint32_t)t;
        for (uint32_t up = p[r]; up != r; up = p[r]) r = up;
        p[t] = r;                                             // flatten; the root is an ancestor, so concurrent climbs stay valid
        atomicAdd(&size[r], 1u);
    }
}

// counters[0] = k-mers that stay active for the next level, counters[1] = k-mers placed in an output component
__global__ void __launch_bounds__(256)
cc_classify_kernel(const uint32_t *__restrict__ vals, uint64_t n, uint8_t *__restrict__ active, const uint32_t *__restrict__ parent,
                   const uint32_t *__restrict__ size, long long b1, long long b2, int thr, uint32_t *__restrict__ label,
                   uint32_t *__restrict__ thr_of, unsigned long long *__restrict__ counters) {
    unsigned long long stay = 0, placed = 0;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        if (!active[t]) continue;
        const uint32_t r = parent[t];
        const long long s = (long long)size[r];
        uint8_t next = 0;
        if (s < b1) {
            // small: skipped (ComponentsBuilder.java:72-74,170-171)
        } else if (s <= b2) {
            label[t] = r; thr_of[t] = (uint32_t)thr; placed++;             // :75-78,172-175
        } else if (cc_short(vals[t]) >= thr + 1) {
            next = 1; stay++;                                               // nextHM, :250-254,257-259
        }
        active[t] = next;
    }
    if (stay) atomicAdd(&counters[0], stay);
    if (placed) atomicAdd(&counters[1], placed);
}

}  // namespace mfkc
