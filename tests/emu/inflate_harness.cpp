// inflate_harness.cpp -- TEST INFRASTRUCTURE: exposes metafast_b200/csrc/fast_inflate.h (the gzip decoder of the host ingest
// path) and its CRC-32 to ctypes, so that tests/test_host.py can compare them with Python's zlib on crafted, corrupted and
// random streams.  Never part of the product.
#include "../../metafast_b200/csrc/fast_inflate.h"

extern "C" long fi_inflate(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t piece, char *err, size_t err_cap, uint64_t *members) {
    mfkc::FastInflate *fi = new mfkc::FastInflate();
    fi->reset(in, n);
    size_t total = 0;
    long rc = 0;
    std::vector<char> buf(piece ? piece : 1);
    for (;;) {
        const long r = fi->read(buf.data(), buf.size());
        if (r < 0) { rc = -1; break; }
        if (r == 0) break;
        if (total + (size_t)r > cap) { rc = -2; break; }
        memcpy(out + total, buf.data(), (size_t)r);
        total += (size_t)r;
    }
    if (err && err_cap) snprintf(err, err_cap, "%s", fi->error().c_str());
    if (members) *members = fi->members();
    delete fi;
    return rc < 0 ? rc : (long)total;
}

extern "C" uint32_t fi_crc32(uint32_t crc, const uint8_t *p, size_t n) { return mfkc::crc32_update(crc, p, n); }

#include "../../metafast_b200/csrc/parallel_inflate.h"

// the multi-threaded decoder (parallel_inflate.h); rc -3 = open() declined (file too small for the given segment size)
extern "C" long pi_inflate(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t piece, int threads, size_t segment_bytes,
                           char *err, size_t err_cap) {
    mfkc::ParallelInflate pi;
    if (!pi.open(in, n, threads, segment_bytes)) return -3;
    size_t total = 0;
    long rc = 0;
    std::vector<char> buf(piece ? piece : 1);
    for (;;) {
        const long r = pi.read(buf.data(), buf.size());
        if (r < 0) { rc = -1; break; }
        if (r == 0) break;
        if (total + (size_t)r > cap) { rc = -2; break; }
        memcpy(out + total, buf.data(), (size_t)r);
        total += (size_t)r;
    }
    if (err && err_cap) snprintf(err, err_cap, "%s", pi.error().c_str());
    return rc < 0 ? rc : (long)total;
}

// the BGZF decoder; rc -3 = open() declined (not BGZF / too small)
extern "C" long bz_inflate(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t piece, int threads, size_t group_bytes,
                           char *err, size_t err_cap) {
    mfkc::BgzfInflate bz;
    if (!bz.open(in, n, threads, group_bytes)) return -3;
    size_t total = 0;
    long rc = 0;
    std::vector<char> buf(piece ? piece : 1);
    for (;;) {
        const long r = bz.read(buf.data(), buf.size());
        if (r < 0) { rc = -1; break; }
        if (r == 0) break;
        if (total + (size_t)r > cap) { rc = -2; break; }
        memcpy(out + total, buf.data(), (size_t)r);
        total += (size_t)r;
    }
    if (err && err_cap) snprintf(err, err_cap, "%s", bz.error().c_str());
    return rc < 0 ? rc : (long)total;
}
