// synth.h -- deterministic "Illumina-like" synthetic read generator shared by the host
// (C++) and the device (CUDA) builds.  Counter-based: read(seed, sample, index) is a pure
// function, so host and device generators agree byte for byte (BASELINE.md section 4,
// SURVEY.md 8d).  Bench / test utility -- not part of the reference's behaviour.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MFKC_HD __host__ __device__ __forceinline__
#else
#define MFKC_HD inline
#endif

#define MFKC_SYNTH_MAX_GENOMES 4096

// Community tables, computed on the host only (doubles), consumed by both generators.
struct mfkc_synth_tables {
    uint64_t seed;
    uint32_t n_genomes;
    uint32_t read_len;
    uint32_t sample;
    uint32_t err_ppm_first, err_ppm_last;
    uint32_t n_read_ppm, poly_tail_ppm;
    uint32_t pad;
    uint64_t genome_off[MFKC_SYNTH_MAX_GENOMES + 1];  // start of genome g in the virtual concatenation
    uint64_t cum_weight[MFKC_SYNTH_MAX_GENOMES];      // inclusive cumulative read-sampling weight, scaled to 2^64
};

MFKC_HD uint64_t mfkc_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

MFKC_HD uint64_t mfkc_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

// 2-bit base (index into "ACGT") of the virtual genome concatenation at position q.
MFKC_HD uint32_t mfkc_synth_genome_base(uint64_t seed, uint64_t q) {
    uint64_t w = mfkc_splitmix64((seed ^ 0x47454E4F4D45ULL) + (q >> 5) * 0xD1342543DE82EF95ULL);
    return (uint32_t)(w >> (2 * (q & 31))) & 3u;
}

struct mfkc_synth_read {
    uint64_t key;        // per-read hash state for the per-base error stream
    uint64_t gpos;       // position of read base 0 on the forward strand (global coordinate)
    uint32_t strand;     // 1 = reverse complement
    uint32_t has_n, n_pos;
    uint32_t tail_len, tail_base;   // tail_len = 0: none
};

MFKC_HD void mfkc_synth_read_header(const mfkc_synth_tables &t, uint64_t index, mfkc_synth_read &r) {
    const uint64_t base = mfkc_splitmix64(t.seed ^ ((uint64_t)t.sample << 40) ^ 0x52454144ULL) + index * 0x9E3779B97F4A7C15ULL;
    const uint64_t r0 = mfkc_splitmix64(base);
    const uint64_t r1 = mfkc_splitmix64(base ^ 0x1111111111111111ULL);
    const uint64_t r2 = mfkc_splitmix64(base ^ 0x2222222222222222ULL);
    const uint64_t r3 = mfkc_splitmix64(base ^ 0x3333333333333333ULL);
    // genome by cumulative weight (binary search)
    uint32_t lo = 0, hi = t.n_genomes - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (r0 <= t.cum_weight[mid]) hi = mid; else lo = mid + 1;
    }
    const uint64_t glen = t.genome_off[lo + 1] - t.genome_off[lo];
    const uint64_t start = mfkc_mulhi64(r1, glen - t.read_len + 1);
    r.key = mfkc_splitmix64(base ^ 0x4444444444444444ULL);
    r.gpos = t.genome_off[lo] + start;
    r.strand = (uint32_t)(r2 & 1);
    r.has_n = ((r2 >> 8) % 1000000ULL) < t.n_read_ppm;
    r.n_pos = (uint32_t)((r2 >> 32) % t.read_len);
    const bool tail = (r3 % 1000000ULL) < t.poly_tail_ppm;
    r.tail_base = (uint32_t)(r3 >> 20) & 1u ? 2u : 0u;       // 'G' or 'A' in "ACGT"
    uint32_t tl = 40 + (uint32_t)((r3 >> 24) % 61);
    if (tl > t.read_len) tl = t.read_len;
    r.tail_len = tail ? tl : 0;
}

// True iff read `index` carries an 'N' (cheap: one hash).
MFKC_HD bool mfkc_synth_read_has_n(const mfkc_synth_tables &t, uint64_t index) {
    const uint64_t base = mfkc_splitmix64(t.seed ^ ((uint64_t)t.sample << 40) ^ 0x52454144ULL) + index * 0x9E3779B97F4A7C15ULL;
    const uint64_t r2 = mfkc_splitmix64(base ^ 0x2222222222222222ULL);
    return ((r2 >> 8) % 1000000ULL) < t.n_read_ppm;
}

// ASCII base j of the read.
MFKC_HD uint8_t mfkc_synth_read_base(const mfkc_synth_tables &t, const mfkc_synth_read &r, uint32_t j) {
    if (r.has_n && j == r.n_pos) return 'N';
    uint32_t x;
    if (r.tail_len && j >= t.read_len - r.tail_len) {
        x = r.tail_base;
    } else {
        const uint64_t p = r.strand ? r.gpos + (t.read_len - 1 - j) : r.gpos + j;
        x = mfkc_synth_genome_base(t.seed, p);
        if (r.strand) x = 3u - x;                       // complement in "ACGT" order
        const uint64_t e = mfkc_splitmix64(r.key + j);
        const uint32_t ppm = t.err_ppm_first +
            (uint32_t)(((uint64_t)(t.err_ppm_last - t.err_ppm_first) * j) / (t.read_len > 1 ? t.read_len - 1 : 1));
        if ((uint32_t)e < ppm * 4295u) x = (x + 1u + (uint32_t)((e >> 32) % 3u)) & 3u;
    }
    return (uint8_t)("ACGT"[x]);
}
