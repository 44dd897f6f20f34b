// fast_inflate.h -- gzip (RFC 1952) / DEFLATE (RFC 1951) decoder for the host ingest path.
//
// The reference reads .gz inputs through java.util.zip.GZIPInputStream inside its synchronized dispatcher
// ([itmo]/io/readers/FastqGZReader.java:25-33, src/io/ReadsDispatcher.java:34-38); libmfkc used zlib's gzread, whose
// inflate (~300 MB/s of text) was the ceiling of every real FASTQ.gz run: the GPU path consumes two orders of magnitude
// more.  This decoder is written for throughput on one core: 64-bit bit buffer refilled with one unaligned load, an
// 11-bit primary table for literals / lengths (8-bit for distances) with sub-tables for longer codes, up to three
// literals per refill, 8-byte match copies.  It streams: `read` hands out decoded text in caller-sized pieces, keeps the
// 32 KiB window itself, walks concatenated members (bgzip, pigz -i, cat a.gz b.gz) and checks CRC-32 and ISIZE of every
// member.  Errors are reported, never papered over; MFKC_INFLATE=zlib selects the old path.
//
// Plain C++17, no CUDA.  Input is a memory range (the reader maps the file).
#pragma once
#include <stdint.h>
#include <string.h>
#include <zlib.h>   // crc32() for short tails and CPUs without PCLMULQDQ
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <string>
#include <vector>

namespace mfkc {

#if defined(__x86_64__)
// reflected CRC-32 (polynomial 0xEDB88320) by carry-less multiplication, 64 bytes per iteration (the folding scheme of
// Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ Instruction", Intel 2009, with the folding
// constants for this polynomial as published there and used by zlib forks -- the routine follows the structure of
// Chromium zlib's crc32_simd.c, (c) The Chromium Authors, BSD-style licence -- ); len >= 64 and a multiple of 16; crc in / out = the raw register
// (the complement of zlib's value).  Checked against zlib's crc32 in tests/test_host.py.
__attribute__((target("pclmul,sse4.1")))
inline uint32_t crc32_clmul_raw(const uint8_t *buf, size_t len, uint32_t crc) {
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596LL, 0x0154442bd4LL);
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009eLL, 0x01751997d0LL);
    const __m128i k5k0 = _mm_set_epi64x(0, 0x0163cd6124LL);
    const __m128i poly = _mm_set_epi64x(0x01f7011641LL, 0x01db710641LL);
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128((const __m128i *)(buf + 0x00)); x2 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
    x3 = _mm_loadu_si128((const __m128i *)(buf + 0x20)); x4 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    x0 = k1k2; buf += 64; len -= 64;
    while (len >= 64) {
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128((const __m128i *)(buf + 0x00)); y6 = _mm_loadu_si128((const __m128i *)(buf + 0x10));
        y7 = _mm_loadu_si128((const __m128i *)(buf + 0x20)); y8 = _mm_loadu_si128((const __m128i *)(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64; len -= 64;
    }
    x0 = k3k4;
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128((const __m128i *)buf);
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16; len -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8); x1 = _mm_xor_si128(x1, x2);
    x0 = k5k0;
    x2 = _mm_srli_si128(x1, 4); x1 = _mm_and_si128(x1, x3); x1 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_xor_si128(x1, x2);
    x0 = poly;
    x2 = _mm_and_si128(x1, x3); x2 = _mm_clmulepi64_si128(x2, x0, 0x10); x2 = _mm_and_si128(x2, x3); x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
#endif

// zlib-compatible crc32 update
inline uint32_t crc32_update(uint32_t crc, const uint8_t *p, size_t n) {
#if defined(__x86_64__)
    static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (have && n >= 64) {
        const size_t body = n & ~(size_t)15;
        crc = ~crc32_clmul_raw(p, body, ~crc);
        p += body; n -= body;
    }
#endif
    while (n) { const size_t part = n > (1u << 30) ? (1u << 30) : n; crc = (uint32_t)crc32(crc, p, (uInt)part); p += part; n -= part; }
    return crc;
}

class FastInflate {
public:
    // `in` must stay valid while reading
    void reset(const uint8_t *in, size_t n) {
        in_ = in; in_next_ = in; in_end_ = in + n;
        bitbuf_ = 0; bitcnt_ = 0;
        state_ = ST_MEMBER_HEADER;
        if (buf_.empty()) buf_.resize(kWindow + kBatch + kSlack);
        out_begin_ = out_next_ = out_read_ = buf_.data() + kWindow;
        hist_ = 0;
        err_.clear(); eof_ = n == 0; members_ = 0;
        if (eof_) state_ = ST_END;
    }
    // gzip magic at the start?  (gzopen reads anything else "transparently" as plain text; the caller does the same)
    static bool looks_like_gzip(const uint8_t *in, size_t n) { return n >= 2 && in[0] == 0x1f && in[1] == 0x8b; }
    const std::string &error() const { return err_; }
    bool failed() const { return !err_.empty(); }
    uint64_t members() const { return members_; }

    // Up to n decoded bytes into dst; 0 = end of data, -1 = error (error())
    long read(char *dst, size_t n) {
        size_t done = 0;
        while (done < n) {
            if (out_read_ == out_next_) {
                if (eof_ || failed()) break;
                if (!decode_more()) break;
                continue;
            }
            const size_t take = std::min<size_t>(n - done, (size_t)(out_next_ - out_read_));
            memcpy(dst + done, out_read_, take);
            out_read_ += take; done += take;
        }
        if (failed() && done == 0) return -1;
        return (long)done;
    }

protected:
    static constexpr size_t kWindow = 32768, kBatch = 1u << 20, kSlack = 512;
    static constexpr int LIT_BITS = 11, DIST_BITS = 8;
    // table entry: value << 16 | kind << 13 | extra_bits << 8 | code bits to consume.  Bit 15 = a literal entry; K_LIT2 =
    // two literals in one primary entry (both codes fit the primary index: value = first | second << 8)
    enum { K_INVALID = 0, K_BASE = 1, K_EOB = 2, K_SUB = 3, K_LITERAL = 4, K_LIT2 = 6 };
    enum State { ST_MEMBER_HEADER, ST_BLOCK_HEADER, ST_STORED, ST_HUFF, ST_TRAILER, ST_END };

    const uint8_t *in_ = nullptr, *in_next_ = nullptr, *in_end_ = nullptr;
    uint64_t bitbuf_ = 0; int bitcnt_ = 0;
    State state_ = ST_END;
    bool final_block_ = false, eof_ = false;
    uint32_t stored_left_ = 0;
    std::vector<uint8_t> buf_;                     // [window | batch | slack]
    uint8_t *out_begin_ = nullptr, *out_next_ = nullptr, *out_read_ = nullptr;
    size_t hist_ = 0;                              // valid bytes of history in front of out_begin_ (this member only)
    uint32_t crc_ = 0; uint32_t isize_ = 0;
    uint64_t members_ = 0;
    std::string err_;
    uint32_t lit_[(1 << LIT_BITS) + 1024], dist_[(1 << DIST_BITS) + 512];

    bool fail(const char *m) { if (err_.empty()) err_ = m; state_ = ST_END; return false; }

    // ---- bit input
    inline void refill() {
        if (in_end_ - in_next_ >= 8) {
            uint64_t w; memcpy(&w, in_next_, 8);
            bitbuf_ |= w << bitcnt_;
            in_next_ += (63 - bitcnt_) >> 3;
            bitcnt_ |= 56;
        } else {
            while (bitcnt_ <= 56 && in_next_ < in_end_) { bitbuf_ |= (uint64_t)*in_next_++ << bitcnt_; bitcnt_ += 8; }
        }
    }
    inline bool need(int n) { if (bitcnt_ < n) { refill(); if (bitcnt_ < n) return false; } return true; }
    inline uint32_t peek(int n) const { return (uint32_t)(bitbuf_ & ((1ull << n) - 1)); }
    inline void drop(int n) { bitbuf_ >>= n; bitcnt_ -= n; }
    // give whole unused bytes back to the input and forget the rest (byte alignment)
    void align_to_byte() {
        if (bitcnt_ < 0) { fail("unexpected end of the gzip stream"); bitbuf_ = 0; bitcnt_ = 0; return; }
        drop(bitcnt_ & 7);
        in_next_ -= bitcnt_ >> 3;
        bitbuf_ = 0; bitcnt_ = 0;
    }

    // ---- canonical Huffman -> lookup table; false: over-subscribed or incomplete (zlib's inflate_table rules)
    static bool build(const uint8_t *lens, int n, uint32_t *table, int primary_bits, int table_cap, const uint16_t *base, const uint8_t *extra,
                      int first_base_sym, bool is_lit) {
        int count[16] = {0};
        for (int i = 0; i < n; i++) count[lens[i]]++;
        int max_len = 15;
        while (max_len > 0 && !count[max_len]) max_len--;
        const int total = 1 << primary_bits;
        if (max_len == 0) {                                   // no codes at all: every lookup is invalid (legal for distances)
            for (int i = 0; i < total; i++) table[i] = (uint32_t)K_INVALID << 13 | 1;
            return true;
        }
        int left = 1;
        for (int len = 1; len <= 15; len++) { left <<= 1; left -= count[len]; if (left < 0) return false; }
        if (left > 0 && max_len != 1) return false;            // incomplete set
        uint16_t next_code[16]; uint32_t code = 0;
        count[0] = 0;
        for (int len = 1; len <= 15; len++) { code = (code + (uint32_t)count[len - 1]) << 1; next_code[len] = (uint16_t)code; }
        for (int i = 0; i < total; i++) table[i] = (uint32_t)K_INVALID << 13 | 1;
        // sub-table sizes: longest code under each primary prefix
        int sub_bits[1 << LIT_BITS];
        if (max_len > primary_bits) {
            memset(sub_bits, 0, sizeof(int) * (size_t)total);
            uint16_t nc[16]; memcpy(nc, next_code, sizeof nc);
            for (int s = 0; s < n; s++) {
                const int len = lens[s];
                if (!len) continue;
                const uint32_t c = nc[len]++;
                if (len > primary_bits) {
                    const uint32_t r = reverse(c, len) & (uint32_t)(total - 1);
                    if (len - primary_bits > sub_bits[r]) sub_bits[r] = len - primary_bits;
                }
            }
            int next_free = total;
            for (int r = 0; r < total; r++)
                if (sub_bits[r]) {
                    if (next_free + (1 << sub_bits[r]) > table_cap) return false;
                    table[r] = (uint32_t)next_free << 16 | (uint32_t)K_SUB << 13 | (uint32_t)sub_bits[r] << 8 | (uint32_t)primary_bits;
                    for (int j = 0; j < (1 << sub_bits[r]); j++) table[next_free + j] = (uint32_t)K_INVALID << 13 | 1;
                    next_free += 1 << sub_bits[r];
                }
        }
        for (int s = 0; s < n; s++) {
            const int len = lens[s];
            if (!len) continue;
            const uint32_t c = next_code[len]++;
            const uint32_t r = reverse(c, len);
            uint32_t e;
            if (is_lit && s < 256) e = (uint32_t)s << 16 | (uint32_t)K_LITERAL << 13;
            else if (is_lit && s == 256) e = (uint32_t)K_EOB << 13;
            else {
                const int b = s - first_base_sym;
                if (b < 0 || (is_lit ? b >= 29 : b >= 30)) e = (uint32_t)K_INVALID << 13;      // length codes 286/287, distance codes 30/31
                else e = (uint32_t)base[b] << 16 | (uint32_t)K_BASE << 13 | (uint32_t)extra[b] << 8;
            }
            if (len <= primary_bits) {
                e |= (uint32_t)len;
                for (uint32_t i = r; i < (uint32_t)total; i += 1u << len) table[i] = e;
            } else {
                const uint32_t p = r & (uint32_t)(total - 1);
                const uint32_t sub = table[p] >> 16; const int sb = (int)((table[p] >> 8) & 31);
                e |= (uint32_t)(len - primary_bits);
                for (uint32_t i = r >> primary_bits; i < (1u << sb); i += 1u << (len - primary_bits)) table[sub + i] = e;
            }
        }
        if (is_lit) {                                          // fuse pairs of short literals (DNA and quality text: 2-5 bit codes)
            uint32_t orig[1 << LIT_BITS];
            memcpy(orig, table, sizeof orig);
            for (int i = 0; i < total; i++) {
                const uint32_t e1 = orig[i];
                if ((e1 >> 13 & 7) != K_LITERAL) continue;
                const int l1 = (int)(e1 & 0xFF);
                if (l1 >= primary_bits) continue;
                const uint32_t e2 = orig[i >> l1];
                const int l2 = (int)(e2 & 0xFF);
                if ((e2 >> 13 & 7) != K_LITERAL || l1 + l2 > primary_bits) continue;
                table[i] = (e2 >> 16 & 0xFF) << 24 | (e1 >> 16 & 0xFF) << 16 | (uint32_t)K_LIT2 << 13 | (uint32_t)(l1 + l2);
            }
        }
        return true;
    }
    static uint32_t reverse(uint32_t c, int len) { uint32_t r = 0; for (int i = 0; i < len; i++) { r = (r << 1) | (c & 1); c >>= 1; } return r; }

    bool read_block_header() {
        static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        if (!need(3)) return fail("unexpected end of the gzip stream");
        final_block_ = peek(1); drop(1);
        const uint32_t type = peek(2); drop(2);
        if (type == 0) {
            align_to_byte();
            if (in_end_ - in_next_ < 4) return fail("unexpected end of the gzip stream");
            const uint32_t len = in_next_[0] | (uint32_t)in_next_[1] << 8, nlen = in_next_[2] | (uint32_t)in_next_[3] << 8;
            in_next_ += 4;
            if ((len ^ 0xFFFFu) != nlen) return fail("invalid stored block lengths");
            stored_left_ = len;
            state_ = ST_STORED;
            return true;
        }
        uint8_t lens[288 + 32];
        if (type == 1) {
            for (int i = 0; i < 144; i++) lens[i] = 8;
            for (int i = 144; i < 256; i++) lens[i] = 9;
            for (int i = 256; i < 280; i++) lens[i] = 7;
            for (int i = 280; i < 288; i++) lens[i] = 8;
            for (int i = 0; i < 32; i++) lens[288 + i] = 5;
            if (!build(lens, 288, lit_, LIT_BITS, (int)(sizeof lit_ / 4), len_base, len_extra, 257, true) ||
                !build(lens + 288, 32, dist_, DIST_BITS, (int)(sizeof dist_ / 4), dist_base, dist_extra, 0, false))
                return fail("internal: fixed tables");
            state_ = ST_HUFF;
            return true;
        }
        if (type == 3) return fail("invalid block type");
        if (!need(14)) return fail("unexpected end of the gzip stream");
        const int nlen = (int)peek(5) + 257; drop(5);
        const int ndist = (int)peek(5) + 1; drop(5);
        const int ncode = (int)peek(4) + 4; drop(4);
        if (nlen > 286 || ndist > 30) return fail("too many length or distance symbols");
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t cl[19] = {0};
        for (int i = 0; i < ncode; i++) { if (!need(3)) return fail("unexpected end of the gzip stream"); cl[order[i]] = (uint8_t)peek(3); drop(3); }
        uint32_t cl_table[128];
        {
            // the code-length code must be complete (zlib: type == CODES never incomplete)
            int count[8] = {0}; for (int i = 0; i < 19; i++) count[cl[i]]++;
            int left = 1; for (int len = 1; len <= 7; len++) { left <<= 1; left -= count[len]; if (left < 0) return fail("invalid code lengths set"); }
            if (left > 0) return fail("invalid code lengths set");
            uint16_t next_code[8]; uint32_t code = 0; count[0] = 0;
            for (int len = 1; len <= 7; len++) { code = (code + (uint32_t)count[len - 1]) << 1; next_code[len] = (uint16_t)code; }
            for (int s = 0; s < 19; s++) {
                const int len = cl[s]; if (!len) continue;
                const uint32_t r = reverse(next_code[len]++, len);
                for (uint32_t i = r; i < 128; i += 1u << len) cl_table[i] = (uint32_t)s << 8 | (uint32_t)len;
            }
        }
        int i = 0;
        while (i < nlen + ndist) {
            if (!need(7 + 7)) { refill(); if (bitcnt_ < 1) return fail("unexpected end of the gzip stream"); }
            const uint32_t e = cl_table[peek(7)];
            const int sym = (int)(e >> 8), len = (int)(e & 0xFF);
            if (bitcnt_ < len) return fail("unexpected end of the gzip stream");
            drop(len);
            if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
            int rep; uint8_t val = 0;
            if (sym == 16) {
                if (i == 0) return fail("invalid bit length repeat");
                val = lens[i - 1];
                if (bitcnt_ < 2) return fail("unexpected end of the gzip stream");
                rep = 3 + (int)peek(2); drop(2);
            } else if (sym == 17) {
                if (bitcnt_ < 3) return fail("unexpected end of the gzip stream");
                rep = 3 + (int)peek(3); drop(3);
            } else {
                if (bitcnt_ < 7) return fail("unexpected end of the gzip stream");
                rep = 11 + (int)peek(7); drop(7);
            }
            if (i + rep > nlen + ndist) return fail("invalid bit length repeat");
            while (rep--) lens[i++] = val;
        }
        if (lens[256] == 0) return fail("invalid code -- missing end-of-block");
        if (!build(lens, nlen, lit_, LIT_BITS, (int)(sizeof lit_ / 4), len_base, len_extra, 257, true)) return fail("invalid literal/lengths set");
        if (!build(lens + nlen, ndist, dist_, DIST_BITS, (int)(sizeof dist_ / 4), dist_base, dist_extra, 0, false)) return fail("invalid distances set");
        state_ = ST_HUFF;
        return true;
    }

    // the hot loop: literals and matches until the block ends or the batch is full
    bool huff_block(uint8_t *out_limit) {
        uint8_t *out = out_next_;
        const uint8_t *const win_start = out_begin_ - hist_;
        const uint32_t lmask = (1u << LIT_BITS) - 1, dmask = (1u << DIST_BITS) - 1;
        while (out < out_limit) {
            refill();
            uint32_t e = lit_[bitbuf_ & lmask];
            if (e & 0x8000u) {                                     // up to three literal entries (1-2 literals each, <= 11 bits) per refill
                drop((int)(e & 0xFF)); out[0] = (uint8_t)(e >> 16); out[1] = (uint8_t)(e >> 24); out += 1 + (e >> 14 & 1);
                e = lit_[bitbuf_ & lmask];
                if (e & 0x8000u) {
                    drop((int)(e & 0xFF)); out[0] = (uint8_t)(e >> 16); out[1] = (uint8_t)(e >> 24); out += 1 + (e >> 14 & 1);
                    e = lit_[bitbuf_ & lmask];
                    if (e & 0x8000u) {
                        drop((int)(e & 0xFF)); out[0] = (uint8_t)(e >> 16); out[1] = (uint8_t)(e >> 24); out += 1 + (e >> 14 & 1);
                        if (bitcnt_ < 0) { out_next_ = out; return fail("unexpected end of the gzip stream"); }
                        continue;
                    }
                }
                refill();
                e = lit_[bitbuf_ & lmask];
            }
            int kind = (int)((e >> 13) & 7);
            if (kind == K_SUB) {
                drop((int)(e & 0xFF));
                e = lit_[(e >> 16) + (uint32_t)(bitbuf_ & ((1u << ((e >> 8) & 31)) - 1))];
                kind = (int)((e >> 13) & 7);
                if (kind == K_LITERAL) {
                    drop((int)(e & 0xFF)); *out++ = (uint8_t)(e >> 16);
                    if (bitcnt_ < 0) { out_next_ = out; return fail("unexpected end of the gzip stream"); }
                    continue;
                }
            }
            if (kind == K_EOB) {
                drop((int)(e & 0xFF));
                if (bitcnt_ < 0) { out_next_ = out; return fail("unexpected end of the gzip stream"); }
                out_next_ = out;
                state_ = final_block_ ? ST_TRAILER : ST_BLOCK_HEADER;
                return true;
            }
            if (kind != K_BASE) { out_next_ = out; return fail(bitcnt_ < 15 ? "unexpected end of the gzip stream" : "invalid literal/length code"); }
            drop((int)(e & 0xFF));
            const int lx = (int)((e >> 8) & 31);
            uint32_t length = (e >> 16) + (uint32_t)(bitbuf_ & ((1u << lx) - 1));
            drop(lx);
            if (bitcnt_ < 32) refill();                            // distance code (15) + extra (13)
            uint32_t d = dist_[bitbuf_ & dmask];
            int dk = (int)((d >> 13) & 7);
            if (dk == K_SUB) {
                drop((int)(d & 0xFF));
                d = dist_[(d >> 16) + (uint32_t)(bitbuf_ & ((1u << ((d >> 8) & 31)) - 1))];
                dk = (int)((d >> 13) & 7);
            }
            if (dk != K_BASE) { out_next_ = out; return fail(bitcnt_ < 15 ? "unexpected end of the gzip stream" : "invalid distance code"); }
            drop((int)(d & 0xFF));
            const int dx = (int)((d >> 8) & 31);
            const uint32_t distance = (d >> 16) + (uint32_t)(bitbuf_ & ((1u << dx) - 1));
            drop(dx);
            if (bitcnt_ < 0) { out_next_ = out; return fail("unexpected end of the gzip stream"); }
            if (distance > (size_t)(out - win_start)) { out_next_ = out; return fail("invalid distance too far back"); }
            const uint8_t *src = out - distance;
            uint8_t *const end = out + length;
            if (distance >= 8) {
                // most matches in DNA / quality text are short: two unconditional words, then a loop for the rest
                uint64_t w0, w1; memcpy(&w0, src, 8); memcpy(out, &w0, 8); memcpy(&w1, src + 8, 8); memcpy(out + 8, &w1, 8);
                if (length > 16) {
                    src += 16; out += 16;
                    do { uint64_t w; memcpy(&w, src, 8); memcpy(out, &w, 8); src += 8; out += 8; } while (out < end);
                }
            } else if (distance == 1) {
                memset(out, *src, length);
            } else {
                do { *out++ = *src++; } while (out < end);
            }
            out = end;
        }
        out_next_ = out;
        return true;
    }

    bool member_header() {
        align_to_byte();
        const uint8_t *p = in_next_;
        const size_t left = (size_t)(in_end_ - p);
        if (members_ > 0 && (left < 2 || p[0] != 0x1f || p[1] != 0x8b)) { eof_ = true; state_ = ST_END; return true; }   // trailing garbage is ignored (gzread)
        if (left < 10) return fail(members_ ? "unexpected end of the gzip stream" : "not a gzip file");
        if (p[0] != 0x1f || p[1] != 0x8b) return fail("not a gzip file");
        if (p[2] != 8) return fail("unknown compression method");
        const int flg = p[3];
        if (flg & 0xE0) return fail("unknown header flags set");
        p += 10;
        auto short_of = [&](size_t n) { return (size_t)(in_end_ - p) < n; };
        if (flg & 4) { if (short_of(2)) return fail("unexpected end of the gzip stream"); const size_t xlen = p[0] | (size_t)p[1] << 8; p += 2; if (short_of(xlen)) return fail("unexpected end of the gzip stream"); p += xlen; }
        for (int bit = 8; bit <= 16; bit <<= 1)
            if (flg & bit) { while (p < in_end_ && *p) p++; if (p >= in_end_) return fail("unexpected end of the gzip stream"); p++; }
        if (flg & 2) {                                                   // FHCRC: the low 16 bits of the CRC-32 of the header so far (RFC 1952)
            if (short_of(2)) return fail("unexpected end of the gzip stream");
            const uint32_t want = p[0] | (uint32_t)p[1] << 8;
            if ((crc32_update(0, in_next_, (size_t)(p - in_next_)) & 0xFFFFu) != want) return fail("gzip header CRC mismatch");
            p += 2;
        }
        in_next_ = p;
        crc_ = 0; isize_ = 0; hist_ = 0;
        members_++;
        state_ = ST_BLOCK_HEADER;
        return true;
    }

    bool trailer() {
        align_to_byte();
        if (in_end_ - in_next_ < 8) return fail("unexpected end of the gzip stream");
        uint32_t crc, isz; memcpy(&crc, in_next_, 4); memcpy(&isz, in_next_ + 4, 4);
        in_next_ += 8;
        if (crc != crc_) return fail("incorrect data check");
        if (isz != isize_) return fail("incorrect length check");
        state_ = ST_MEMBER_HEADER;
        return true;
    }

    // account for freshly decoded bytes [from, out_next_) of the current member
    void account(const uint8_t *from) {
        const size_t n = (size_t)(out_next_ - from);
        if (!n) return;
        crc_ = crc32_update(crc_, from, n);
        isize_ += (uint32_t)n;
    }

    // decode until at least one byte is available or the data ends; false = nothing more (end or error)
    bool decode_more() {
        // slide: keep the last 32 KiB of this member in front of the batch area
        const size_t produced = (size_t)(out_next_ - out_begin_);
        if (state_ == ST_MEMBER_HEADER) hist_ = 0;             // a new member starts with an empty window
        else if (produced) {
            const size_t keep = std::min<size_t>(kWindow, hist_ + produced);
            memmove(out_begin_ - keep, out_next_ - keep, keep);
            hist_ = keep;
        }
        out_next_ = out_read_ = out_begin_;
        uint8_t *const limit = out_begin_ + kBatch;
        const uint8_t *from = out_next_;
        while (out_next_ < limit && !failed()) {
            switch (state_) {
            case ST_MEMBER_HEADER:
                if (in_next_ >= in_end_ && bitcnt_ < 8 && members_ > 0) { eof_ = true; state_ = ST_END; break; }
                member_header();
                from = out_next_;
                break;
            case ST_BLOCK_HEADER: read_block_header(); break;
            case ST_STORED: {
                const size_t take = std::min<size_t>(std::min<size_t>(stored_left_, (size_t)(limit - out_next_)), (size_t)(in_end_ - in_next_));
                memcpy(out_next_, in_next_, take);
                out_next_ += take; in_next_ += take; stored_left_ -= (uint32_t)take;
                if (stored_left_ && in_next_ >= in_end_) { fail("unexpected end of the gzip stream"); break; }
                if (!stored_left_) state_ = final_block_ ? ST_TRAILER : ST_BLOCK_HEADER;
                break;
            }
            case ST_HUFF: huff_block(limit); break;
            case ST_TRAILER:
                account(from); from = out_next_;
                trailer();
                // the next member starts in a fresh batch (empty window: its matches must not reach into this member)
                if (!failed() && out_next_ != out_begin_) goto out;
                break;
            case ST_END: goto out;
            }
            if (state_ == ST_END) break;
        }
    out:
        account(from);
        return out_next_ != out_read_;
    }
};

}  // namespace mfkc
