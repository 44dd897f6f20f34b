// radix_sort.cuh -- hand-written LSD radix sort for 64-bit keys with an optional payload
// (sm_100a).  Used by K4 (order the emitted records by key), by the sort-and-run-length
// counting variant K3' and by the deterministic shard merge.
//
// One pass = 8-bit digit.  Work unit = one WARP owning a contiguous sub-tile of RS_WARP_ITEMS
// keys and a private 256-entry cursor array in shared memory:
//   rs_hist_kernel     per-warp digit histogram                      hist[part][digit]
//   rs_chunk_kernel    column sums over chunks of RS_CHUNK parts      chunk_tot[chunk][digit]
//   rs_base_kernel     exclusive scan in (digit, chunk) order         chunk_tot -> chunk base
//   rs_offsets_kernel  hist[part][digit] -> global start offset of (part, digit)
//   rs_scatter_kernel  stable scatter: lanes rank themselves inside the warp with match.any,
//                      the warp's cursors give the global position -- no block-level sync and
//                      no second read of the tile.
// HBM traffic per pass: 2 reads + 1 write of the keys (+ payload once each way); the write
// frontier (parts x 256 x 32-byte sectors) stays L2-resident, so partial-sector stores merge
// in L2 before reaching DRAM.
#pragma once
#include "device_common.cuh"

namespace mfkc {

constexpr int RS_RADIX = 256;
constexpr int RS_WARPS = 8;                 // warps per block
constexpr int RS_WARP_ITEMS = 4096;         // keys per warp sub-tile (128 rounds of 32)
constexpr int RS_CHUNK = 128;               // parts per scan chunk

__global__ void __launch_bounds__(RS_WARPS * 32)
rs_hist_kernel(const unsigned long long *__restrict__ keys, uint64_t n, int shift, uint32_t n_parts,
               uint32_t *__restrict__ hist) {
    __shared__ uint32_t s_cnt[RS_WARPS][RS_RADIX];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t part = blockIdx.x * RS_WARPS + warp; part < n_parts; part += gridDim.x * RS_WARPS) {
        for (int d = lane; d < RS_RADIX; d += 32) s_cnt[warp][d] = 0;
        __syncwarp();
        const uint64_t lo = (uint64_t)part * RS_WARP_ITEMS;
        const uint64_t hi = lo + RS_WARP_ITEMS < n ? lo + RS_WARP_ITEMS : n;
        for (uint64_t i = lo + lane; i < hi; i += 32) {
            const uint32_t d = (uint32_t)(keys[i] >> shift) & (RS_RADIX - 1);
            atomicAdd(&s_cnt[warp][d], 1u);
        }
        __syncwarp();
        for (int d = lane; d < RS_RADIX; d += 32) hist[(uint64_t)part * RS_RADIX + d] = s_cnt[warp][d];
        __syncwarp();
    }
}

// block b sums hist over parts [b*RS_CHUNK, (b+1)*RS_CHUNK): thread d owns digit d
__global__ void __launch_bounds__(RS_RADIX)
rs_chunk_kernel(const uint32_t *__restrict__ hist, uint32_t n_parts, unsigned long long *__restrict__ chunk_tot) {
    const uint32_t d = threadIdx.x;
    const uint32_t p0 = blockIdx.x * RS_CHUNK;
    const uint32_t p1 = p0 + RS_CHUNK < n_parts ? p0 + RS_CHUNK : n_parts;
    unsigned long long s = 0;
    for (uint32_t p = p0; p < p1; p++) s += hist[(uint64_t)p * RS_RADIX + d];
    chunk_tot[(uint64_t)blockIdx.x * RS_RADIX + d] = s;
}

// single block: exclusive scan of chunk_tot in digit-major order.  Thread d first totals its
// digit column, an in-block scan over the 256 digit totals gives the digit base, then the
// thread walks its column again.  Also reports whether one digit holds all n keys (pass can
// be skipped).
__global__ void __launch_bounds__(RS_RADIX)
rs_base_kernel(unsigned long long *__restrict__ chunk_tot, uint32_t n_chunks, uint64_t n, uint32_t *__restrict__ trivial) {
    __shared__ unsigned long long s_tot[RS_RADIX];
    const uint32_t d = threadIdx.x;
    unsigned long long tot = 0;
    for (uint32_t c = 0; c < n_chunks; c++) tot += chunk_tot[(uint64_t)c * RS_RADIX + d];
    s_tot[d] = tot;
    __syncthreads();
    if (d == 0) {
        unsigned long long run = 0; uint32_t triv = 0;
        for (int i = 0; i < RS_RADIX; i++) { const unsigned long long t = s_tot[i]; if (t == n) triv = 1; s_tot[i] = run; run += t; }
        *trivial = triv;
    }
    __syncthreads();
    unsigned long long run = s_tot[d];
    for (uint32_t c = 0; c < n_chunks; c++) {
        const unsigned long long t = chunk_tot[(uint64_t)c * RS_RADIX + d];
        chunk_tot[(uint64_t)c * RS_RADIX + d] = run;
        run += t;
    }
}

// hist[part][d] (counts) -> offs[part][d] (global start offsets, 64-bit)
__global__ void __launch_bounds__(RS_RADIX)
rs_offsets_kernel(const uint32_t *__restrict__ hist, uint32_t n_parts, const unsigned long long *__restrict__ chunk_base,
                  unsigned long long *__restrict__ offs) {
    const uint32_t d = threadIdx.x;
    const uint32_t p0 = blockIdx.x * RS_CHUNK;
    const uint32_t p1 = p0 + RS_CHUNK < n_parts ? p0 + RS_CHUNK : n_parts;
    unsigned long long run = chunk_base[(uint64_t)blockIdx.x * RS_RADIX + d];
    for (uint32_t p = p0; p < p1; p++) {
        const uint32_t c = hist[(uint64_t)p * RS_RADIX + d];
        offs[(uint64_t)p * RS_RADIX + d] = run;
        run += c;
    }
}

template <typename V, bool HAS_V>
__global__ void __launch_bounds__(RS_WARPS * 32)
rs_scatter_kernel(const unsigned long long *__restrict__ keys_in, const V *__restrict__ vals_in, uint64_t n, int shift,
                  uint32_t n_parts, const unsigned long long *__restrict__ offs,
                  unsigned long long *__restrict__ keys_out, V *__restrict__ vals_out) {
    __shared__ unsigned long long s_cur[RS_WARPS][RS_RADIX];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = lanemask_lt();
    for (uint32_t part = blockIdx.x * RS_WARPS + warp; part < n_parts; part += gridDim.x * RS_WARPS) {
        for (int d = lane; d < RS_RADIX; d += 32) s_cur[warp][d] = offs[(uint64_t)part * RS_RADIX + d];
        __syncwarp();
        const uint64_t lo = (uint64_t)part * RS_WARP_ITEMS;
        const uint64_t hi = lo + RS_WARP_ITEMS < n ? lo + RS_WARP_ITEMS : n;
        for (uint64_t base = lo; base < hi; base += 32) {
            const uint64_t i = base + lane;
            const bool ok = i < hi;
            unsigned long long key = 0; V val = V();
            if (ok) { key = keys_in[i]; if (HAS_V) val = vals_in[i]; }
            const uint32_t d = ok ? ((uint32_t)(key >> shift) & (RS_RADIX - 1)) : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            if (ok) {
                const unsigned long long pos = s_cur[warp][d] + __popc(peers & lt);
                keys_out[pos] = key;
                if (HAS_V) vals_out[pos] = val;
            }
            __syncwarp();
            if (ok && (peers & lt) == 0) s_cur[warp][d] += __popc(peers);    // lowest peer advances the cursor
            __syncwarp();
        }
    }
}

// Workspace sizes for n keys.
struct RadixPlan {
    uint32_t n_parts, n_chunks;
    size_t hist_bytes, offs_bytes, chunk_bytes;
};
inline RadixPlan radix_plan(uint64_t n) {
    RadixPlan p;
    p.n_parts = (uint32_t)((n + RS_WARP_ITEMS - 1) / RS_WARP_ITEMS);
    if (p.n_parts == 0) p.n_parts = 1;
    p.n_chunks = (p.n_parts + RS_CHUNK - 1) / RS_CHUNK;
    p.hist_bytes = (size_t)p.n_parts * RS_RADIX * sizeof(uint32_t);
    p.offs_bytes = (size_t)p.n_parts * RS_RADIX * sizeof(unsigned long long);
    p.chunk_bytes = (size_t)p.n_chunks * RS_RADIX * sizeof(unsigned long long);
    return p;
}

}  // namespace mfkc
