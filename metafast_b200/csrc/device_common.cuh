// device_common.cuh -- shared device helpers for libmfkc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mfkc {

constexpr unsigned long long EMPTY_KEY = ~0ull;   // legal keys are < 2^62 (k <= 31); key 0 (poly-A) is legal
constexpr uint32_t MAX_COUNT = 32767u;            // Short.MAX_VALUE, [itmo]/utils/NumUtils.java:21-26

// One table slot = half a 32-byte DRAM sector: key and count always share a sector, so an
// upsert touches exactly one sector (read + eventual write-back = the 64 B/k-mer of SURVEY 8d).
struct __align__(16) Slot {
    unsigned long long key;
    uint32_t count;      // true count while < 32767; may transiently overshoot under races, clamped on emit
    uint32_t pad;
};
static_assert(sizeof(Slot) == 16, "slot must be 16 bytes");

// 64-bit accumulator slot of the features-calculator set (BigLong2LongHashMap analogue).
struct __align__(16) FcSlot {
    unsigned long long key;
    long long acc;
};

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t k) {   // MurmurHash3 fmix64
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

__host__ __device__ __forceinline__ uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// Home slot of a key in a table of `cap` slots (any capacity, not only powers of two).
__host__ __device__ __forceinline__ uint64_t home_slot(uint64_t key, uint64_t cap) {
    return mulhi64(mix64(key), cap);
}

// Owner shard for hash-range partitioning across GPUs: a hash independent of home_slot().
__host__ __device__ __forceinline__ uint32_t owner_shard(uint64_t key, uint32_t n_shards) {
    return (uint32_t)mulhi64(mix64(key ^ 0x5bd1e9955bd1e995ULL), (uint64_t)n_shards);
}

#if defined(__CUDACC__)
__device__ __forceinline__ ulonglong2 ld_cg_u64x2(const void *p) {   // L1-bypassing 128-bit load
    ulonglong2 v;
    asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ uint4 ld_nc_u128(const void *p) {          // streaming 128-bit load
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ASCII (AaCcGgTt) x4 -> four 2-bit codes A0 G1 C2 T3 ([itmo]/dna/DnaTools.java:46-60),
// first (lowest-address) base in the most significant pair of the returned byte.
//   bit2,bit1 of the ASCII code: A 00, C 01, T 10, G 11  ->  code = ((b2^b1)<<1) | b2
__device__ __forceinline__ uint32_t pack4(uint32_t x) {
    const uint32_t b1 = (x >> 1) & 0x01010101u;
    const uint32_t b2 = (x >> 2) & 0x01010101u;
    const uint32_t c = ((b1 ^ b2) << 1) | b2;          // one code per byte
    return (c * 0x40100401u) >> 24;                    // gather: byte0 -> bits 7:6 ... byte3 -> bits 1:0
}
// Non-zero byte lanes where the byte is NOT one of AaCcGgTt.  The letter is rebuilt from its own bits 1 and 2
// (A 41, C 43, G 47; T would come out as 45 and is patched by 0x11) and compared with the upper-cased input: only
// the four letters are fixed points.  9 instructions per 4 bases (four SIMD byte compares cost 45 on sm_100).
__device__ __forceinline__ uint32_t bad4(uint32_t x) {
    const uint32_t m = x & 0xDFDFDFDFu;
    const uint32_t b1 = (m >> 1) & 0x01010101u, b2 = (m >> 2) & 0x01010101u;
    const uint32_t t = b2 & ~b1;
    const uint32_t expect = 0x41414141u | (b1 << 1) | (b2 << 2);
    return (m ^ expect) ^ (t * 0x11u);
}
// 16 ASCII bases -> one 32-bit word, base 0 in bits 31:30.
__device__ __forceinline__ uint32_t pack16(uint4 v) {
    return (pack4(v.x) << 24) | (pack4(v.y) << 16) | (pack4(v.z) << 8) | pack4(v.w);
}
#endif

}  // namespace mfkc
