// radix_sort.cuh -- hand-written stable LSD radix sort for 64-bit keys with an optional payload
// (sm_100a).  Used by K4 (order the emitted records by key), by the sort-and-run-length
// counting variant K3' and by the merge of sorted (key,count) runs.
//
// One pass = one 8-bit digit over tiles of RS_TILE keys (one CTA per tile):
//   rs_hist_kernel     per-tile digit histogram                       hist[tile][digit]   (u32)
//   rs_chunk_kernel    column sums over chunks of RS_CHUNK tiles      chunk_tot[chunk][digit]
//   rs_base_kernel     exclusive scan in (digit, chunk) order         chunk_tot -> chunk base
//   rs_offsets_kernel  hist[tile][digit] -> global start offset of (tile, digit)
//   rs_scatter_kernel  stable scatter.  Lanes rank themselves inside their warp with match.any,
//                      per-warp digit counters are scanned across the 8 warps, the tile is
//                      reordered by digit in shared memory, and every digit run leaves the CTA as
//                      one contiguous (coalesced, whole-sector) burst.
// The first version of this file scattered straight from registers with per-warp cursors; ncu
// showed DRAM-level read-modify-write of partially written sectors (the open write frontier of
// ~9.5 k warps x 256 digits exceeded L2), 67 ms for 120 M records.  Reordering through shared
// memory removes that.
// HBM traffic per pass: keys read twice (histogram, scatter) + written once, payload once each way.
#pragma once
#include "device_common.cuh"

namespace mfkc {

constexpr int RS_RADIX = 256;
constexpr int RS_THREADS = 256;             // = RS_RADIX: thread d owns digit d in the block-level scans
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;      // 4096 keys per CTA tile
constexpr int RS_WARP_ITEMS = RS_TILE / RS_WARPS;   // 512 consecutive keys per warp
constexpr int RS_CHUNK = 128;               // tiles per scan chunk

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const unsigned long long *__restrict__ keys, uint64_t n, int shift, uint32_t n_tiles,
               uint32_t *__restrict__ hist) {
    __shared__ uint32_t s_cnt[RS_RADIX];
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const uint64_t lo = (uint64_t)tile * RS_TILE;
        unsigned long long k[RS_ITEMS];
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++) {
            const uint64_t i = lo + (uint64_t)r * RS_THREADS + threadIdx.x;
            k[r] = i < n ? keys[i] : ~0ull;
        }
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++) {
            const uint64_t i = lo + (uint64_t)r * RS_THREADS + threadIdx.x;
            if (i < n) atomicAdd(&s_cnt[(uint32_t)(k[r] >> shift) & (RS_RADIX - 1)], 1u);
        }
        __syncthreads();
        hist[(uint64_t)tile * RS_RADIX + threadIdx.x] = s_cnt[threadIdx.x];
        __syncthreads();
    }
}

// block b sums hist over tiles [b*RS_CHUNK, (b+1)*RS_CHUNK): thread d owns digit d
__global__ void __launch_bounds__(RS_RADIX)
rs_chunk_kernel(const uint32_t *__restrict__ hist, uint32_t n_tiles, unsigned long long *__restrict__ chunk_tot) {
    const uint32_t d = threadIdx.x;
    const uint32_t p0 = blockIdx.x * RS_CHUNK;
    const uint32_t p1 = p0 + RS_CHUNK < n_tiles ? p0 + RS_CHUNK : n_tiles;
    unsigned long long s = 0;
#pragma unroll 8
    for (uint32_t p = p0; p < p1; p++) s += hist[(uint64_t)p * RS_RADIX + d];
    chunk_tot[(uint64_t)blockIdx.x * RS_RADIX + d] = s;
}

// single block: exclusive scan of chunk_tot in digit-major order.  Also reports whether one digit
// holds all n keys (the pass can then be skipped: order unchanged).
__global__ void __launch_bounds__(RS_RADIX)
rs_base_kernel(unsigned long long *__restrict__ chunk_tot, uint32_t n_chunks, uint64_t n, uint32_t *__restrict__ trivial) {
    __shared__ unsigned long long s_tot[RS_RADIX];
    const uint32_t d = threadIdx.x;
    unsigned long long tot = 0;
#pragma unroll 8
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
#pragma unroll 8
    for (uint32_t c = 0; c < n_chunks; c++) {
        const unsigned long long t = chunk_tot[(uint64_t)c * RS_RADIX + d];
        chunk_tot[(uint64_t)c * RS_RADIX + d] = run;
        run += t;
    }
}

// hist[tile][d] (counts) -> offs[tile][d] (global start offsets)
__global__ void __launch_bounds__(RS_RADIX)
rs_offsets_kernel(const uint32_t *__restrict__ hist, uint32_t n_tiles, const unsigned long long *__restrict__ chunk_base,
                  unsigned long long *__restrict__ offs) {
    const uint32_t d = threadIdx.x;
    const uint32_t p0 = blockIdx.x * RS_CHUNK;
    const uint32_t p1 = p0 + RS_CHUNK < n_tiles ? p0 + RS_CHUNK : n_tiles;
    unsigned long long run = chunk_base[(uint64_t)blockIdx.x * RS_RADIX + d];
#pragma unroll 8
    for (uint32_t p = p0; p < p1; p++) {
        const uint32_t c = hist[(uint64_t)p * RS_RADIX + d];
        offs[(uint64_t)p * RS_RADIX + d] = run;
        run += c;
    }
}

// MINB = CTAs per SM the kernel is compiled for: 2 (128 registers, no spills) or 3 (80 registers, a few spilled values):
// the kernel is latency-bound (ncu: 23 % issue utilisation at 25 % occupancy), see radix_sort() for which one is used
template <typename V, bool HAS_V, int MINB>
__global__ void __launch_bounds__(RS_THREADS, MINB)
rs_scatter_kernel(const unsigned long long *__restrict__ keys_in, const V *__restrict__ vals_in, uint64_t n, int shift,
                  uint32_t n_tiles, const unsigned long long *__restrict__ offs,
                  unsigned long long *__restrict__ keys_out, V *__restrict__ vals_out) {
    extern __shared__ __align__(16) unsigned char rs_dyn_smem[];    // RS_TILE keys, then RS_TILE payloads
    unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(rs_dyn_smem);
    V *s_vals = reinterpret_cast<V *>(rs_dyn_smem + (size_t)RS_TILE * sizeof(unsigned long long));
    __shared__ uint32_t s_cnt[RS_WARPS][RS_RADIX];       // per-warp digit counts -> exclusive prefix over warps
    __shared__ uint32_t s_dstart[RS_RADIX];              // first tile-local position of digit d
    __shared__ unsigned long long s_goff[RS_RADIX];      // global position of the tile's first key with digit d
    __shared__ uint32_t s_wsum[RS_WARPS];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t lt = lanemask_lt();

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t lo = (uint64_t)tile * RS_TILE;
        const uint32_t tile_n = (uint32_t)((lo + RS_TILE <= n) ? RS_TILE : (n - lo));
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) s_cnt[w][tid] = 0;
        __syncthreads();

        // warp w owns tile positions [w*512, (w+1)*512): round r, lane l -> w*512 + r*32 + l
        // MINB <= 3: keys and payloads stay in registers between ranking and reordering (128 / 80 registers).
        // MINB == 4: only (digit | rank << 8) is kept; the reorder phase reads the tile again (L2 hits): ~56 registers,
        //            4 CTAs per SM to hide the latencies this kernel is bound by.
        constexpr bool RELOAD = MINB >= 4;
        unsigned long long key[RELOAD ? 1 : RS_ITEMS];
        V val[(HAS_V && !RELOAD) ? RS_ITEMS : 1];
        uint32_t rk[RELOAD ? RS_ITEMS : RS_ITEMS / 2];    // RELOAD: digit | rank << 8 per item; else rank inside the warp's digit group, 16 bits each
        if constexpr (!RELOAD) {
#pragma unroll
            for (int r = 0; r < RS_ITEMS; r++) {
                const uint32_t p = warp * RS_WARP_ITEMS + r * 32 + lane;
                key[r] = p < tile_n ? keys_in[lo + p] : ~0ull;
                if (HAS_V) val[r] = p < tile_n ? vals_in[lo + p] : V();
            }
        }
#pragma unroll
        for (int half = 0; half < 2; half++) {
            unsigned long long kh[RELOAD ? RS_ITEMS / 2 : 1];
            if constexpr (RELOAD) {
#pragma unroll
                for (int q = 0; q < RS_ITEMS / 2; q++) {
                    const uint32_t p = warp * RS_WARP_ITEMS + (half * (RS_ITEMS / 2) + q) * 32 + lane;
                    kh[q] = p < tile_n ? keys_in[lo + p] : ~0ull;
                }
            }
#pragma unroll
            for (int q = 0; q < RS_ITEMS / 2; q++) {
                const int r = half * (RS_ITEMS / 2) + q;
                const uint32_t p = warp * RS_WARP_ITEMS + r * 32 + lane;
                const bool ok = p < tile_n;
                const unsigned long long kk = RELOAD ? kh[RELOAD ? q : 0] : key[RELOAD ? 0 : r];
                const uint32_t d = ok ? ((uint32_t)(kk >> shift) & (RS_RADIX - 1)) : 0xFFFFFFFFu;
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                uint32_t rank = 0;
                if (ok) rank = s_cnt[warp][d] + __popc(peers & lt);
                __syncwarp();
                if (ok && (peers & lt) == 0) s_cnt[warp][d] += __popc(peers);     // lowest peer advances the counter
                __syncwarp();
                if constexpr (RELOAD) rk[r] = (d & 0xFFu) | (rank << 8);
                else { if (r & 1) rk[r >> 1] |= rank << 16; else rk[r >> 1] = rank; }
            }
        }
        __syncthreads();

        // thread d: exclusive prefix of digit d over the warps, then exclusive scan over digits
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { const uint32_t t = s_cnt[w][tid]; s_cnt[w][tid] = total; total += t; }
        uint32_t incl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) if (w < (int)warp) wbase += s_wsum[w];
        s_dstart[tid] = wbase + incl - total;
        s_goff[tid] = offs[(uint64_t)tile * RS_RADIX + tid];
        __syncthreads();

        // reorder the tile by digit in shared memory (stable)
        if constexpr (RELOAD) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                unsigned long long kh[RS_ITEMS / 2]; V vh[HAS_V ? RS_ITEMS / 2 : 1];
#pragma unroll
                for (int q = 0; q < RS_ITEMS / 2; q++) {
                    const uint32_t p = warp * RS_WARP_ITEMS + (half * (RS_ITEMS / 2) + q) * 32 + lane;
                    kh[q] = p < tile_n ? keys_in[lo + p] : ~0ull;
                    if (HAS_V) vh[q] = p < tile_n ? vals_in[lo + p] : V();
                }
#pragma unroll
                for (int q = 0; q < RS_ITEMS / 2; q++) {
                    const int r = half * (RS_ITEMS / 2) + q;
                    const uint32_t p = warp * RS_WARP_ITEMS + r * 32 + lane;
                    if (p < tile_n) {
                        const uint32_t d = rk[r] & 0xFFu, rank = rk[r] >> 8;
                        const uint32_t pos = s_dstart[d] + s_cnt[warp][d] + rank;
                        s_keys[pos] = kh[q];
                        if (HAS_V) s_vals[pos] = vh[q];
                    }
                }
            }
        } else {
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++) {
            const uint32_t p = warp * RS_WARP_ITEMS + r * 32 + lane;
            if (p < tile_n) {
                const uint32_t d = (uint32_t)(key[r] >> shift) & (RS_RADIX - 1);
                const uint32_t rank = (r & 1) ? (rk[r >> 1] >> 16) : (rk[r >> 1] & 0xFFFFu);
                const uint32_t pos = s_dstart[d] + s_cnt[warp][d] + rank;
                s_keys[pos] = key[r];
                if (HAS_V) s_vals[pos] = val[r];
            }
        }
        }
        __syncthreads();

        // coalesced copy-out: consecutive threads, consecutive positions of a digit run
        for (uint32_t p = tid; p < tile_n; p += RS_THREADS) {
            const unsigned long long k2 = s_keys[p];
            const uint32_t d = (uint32_t)(k2 >> shift) & (RS_RADIX - 1);
            const unsigned long long g = s_goff[d] + (p - s_dstart[d]);
            keys_out[g] = k2;
            if (HAS_V) vals_out[g] = s_vals[p];
        }
        __syncthreads();
    }
}

template <typename V, bool HAS_V>
constexpr size_t rs_scatter_smem_bytes() { return (size_t)RS_TILE * (sizeof(unsigned long long) + (HAS_V ? sizeof(V) : 0)); }

// Workspace sizes for n keys.
struct RadixPlan {
    uint32_t n_parts, n_chunks;
    size_t hist_bytes, offs_bytes, chunk_bytes;
};
inline RadixPlan radix_plan(uint64_t n) {
    RadixPlan p;
    p.n_parts = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    if (p.n_parts == 0) p.n_parts = 1;
    p.n_chunks = (p.n_parts + RS_CHUNK - 1) / RS_CHUNK;
    p.hist_bytes = (size_t)p.n_parts * RS_RADIX * sizeof(uint32_t);
    p.offs_bytes = (size_t)p.n_parts * RS_RADIX * sizeof(unsigned long long);
    p.chunk_bytes = (size_t)p.n_chunks * RS_RADIX * sizeof(unsigned long long);
    return p;
}

}  // namespace mfkc
