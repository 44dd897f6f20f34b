// parallel_inflate.h -- one gzip member decoded by several threads (host ingest).
//
// A DEFLATE stream has no index, and every block may copy from the 32 KiB of text in front of it, so a single stream is
// normally decoded by one core (fast_inflate.h: ~0.5 GB/s of text; the parsers and the GPU behind it take several
// GB/s).  Here the compressed file is cut into segments and every segment is decoded at the same time:
//   * a thread looks for the first block start in its segment by trying every bit position: a dynamic-Huffman header
//     with complete codes, whose block decodes to text characters only and is followed by another valid header;
//   * from there it decodes with an UNKNOWN window: the output is 16-bit symbols, 0..255 = a byte, 256 + k = "byte k of
//     the 32 KiB in front of my start".  Copies move markers like any other symbol;
//   * the thread of the previous segment decodes on until it arrives -- at a block boundary -- at exactly the bit
//     position the next thread started from.  That meeting is the proof: the predecessor is in sync with the true
//     stream, so a block does start there, and everything the successor decoded is exact up to its markers.  A start
//     that is passed without being met was a false positive: that run is thrown away and the predecessor carries on;
//   * runs are resolved in order: the last 32 KiB of resolved text of one run is the window of the next, markers are
//     replaced (a SIMD narrowing pass with a scalar fix-up where markers remain), CRC-32 per piece, combined at the end
//     and checked against the member trailer together with ISIZE.
// So the result is exact by construction, not by likelihood; the text check only keeps false starts rare.  `read` has
// the interface of FastInflate::read.  A file with several members (bgzip, cat a.gz b.gz) leaves parallel mode at the
// end of its first member and continues with FastInflate.  (The idea of decoding with an unknown window and resolving
// later is that of pugz, Kerbiriou & Chikhi 2019; this implementation is independent.)
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "fast_inflate.h"

namespace mfkc {

class MarkerInflate : public FastInflate {
public:
    static constexpr size_t WIN = 32768;
    static constexpr uint16_t NOTHING = 0xFFFF;               // "in front of the member's first byte": resolving it is an error

    void start_at(const uint8_t *in, size_t n, uint64_t bitpos) {
        in_ = in; in_end_ = in + n; in_next_ = in + (bitpos >> 3);
        bitbuf_ = 0; bitcnt_ = 0; err_.clear(); state_ = ST_BLOCK_HEADER; final_block_ = false; eof_ = false;
        refill();
        drop((int)(bitpos & 7));
    }
    uint64_t bitpos() const { return (uint64_t)(in_next_ - in_) * 8 - (uint64_t)bitcnt_; }
    bool at_final_block_end() const { return state_ == ST_TRAILER; }
    bool at_block_header() const { return state_ == ST_BLOCK_HEADER; }
    using FastInflate::failed;
    using FastInflate::error;

    // skips the gzip member header at the start of `in`; the deflate data starts at bitpos()
    bool skip_member_header(const uint8_t *in, size_t n) {
        in_ = in; in_next_ = in; in_end_ = in + n; bitbuf_ = 0; bitcnt_ = 0; err_.clear(); members_ = 0;
        return member_header() && !failed();
    }
    // after the final block: CRC-32 and ISIZE of the trailer, and where the next member (if any) would start
    bool read_trailer(uint32_t *crc, uint32_t *isize, size_t *next_offset) {
        align_to_byte();
        if (failed() || in_end_ - in_next_ < 8) return false;
        memcpy(crc, in_next_, 4); memcpy(isize, in_next_ + 4, 4);
        *next_offset = (size_t)(in_next_ + 8 - in_);
        return true;
    }

    // Is there a plausible block start at `bitpos`?  (dynamic block, not final, complete codes; the block decodes to text
    // characters and ends; another valid block header follows)
    bool probe(const uint8_t *in, size_t n, uint64_t bitpos) {
        start_at(in, n, bitpos);
        if (bitcnt_ < 17) return false;
        const uint32_t b = peek(17);
        if ((b & 1) || ((b >> 1) & 3) != 2 || ((b >> 3) & 31) > 29 || ((b >> 8) & 31) > 29) return false;
        if (!read_block_header() || state_ != ST_HUFF) return false;
        if (!probe_block()) return false;
        if (state_ != ST_BLOCK_HEADER) return false;
        if (!need(3)) return false;
        const uint32_t type = (peek(3) >> 1) & 3;
        if (type == 3) return false;
        if (type == 2) return read_block_header();
        if (type == 0) {
            drop(3);
            align_to_byte();
            if (failed() || in_end_ - in_next_ < 4) return false;
            const uint32_t len = in_next_[0] | (uint32_t)in_next_[1] << 8, nlen = in_next_[2] | (uint32_t)in_next_[3] << 8;
            return (len ^ 0xFFFFu) == nlen;
        }
        return true;
    }

    // Decodes blocks into 16-bit symbols behind `out` (which has WIN symbols of prefix in front of the piece's first
    // symbol) until the output reaches `limit`, a block ends (returns at every block boundary so that the caller can
    // compare positions) or the final block ends.  false = error.
    bool decode_some(uint16_t *&out, uint16_t *limit) {
        if (state_ == ST_BLOCK_HEADER) { if (!read_block_header()) return false; }
        if (state_ == ST_STORED) {
            size_t take = std::min<size_t>(std::min<size_t>(stored_left_, (size_t)(limit - out)), (size_t)(in_end_ - in_next_));
            for (size_t i = 0; i < take; i++) out[i] = in_next_[i];
            out += take; in_next_ += take; stored_left_ -= (uint32_t)take;
            if (stored_left_ && in_next_ >= in_end_) return fail("unexpected end of the gzip stream");
            if (!stored_left_) state_ = final_block_ ? ST_TRAILER : ST_BLOCK_HEADER;
            return true;
        }
        if (state_ == ST_HUFF) return huff_block16(out, limit);
        return !failed();
    }

private:
    bool probe_block() {
        static const struct Text { bool ok[256]; Text() { for (int c = 0; c < 256; c++) ok[c] = c == 9 || c == 10 || c == 13 || (c >= 32 && c <= 126); } } text;
        const uint32_t lmask = (1u << LIT_BITS) - 1, dmask = (1u << DIST_BITS) - 1;
        for (uint32_t nsym = 0; nsym < (1u << 24); nsym++) {
            refill();
            uint32_t e = lit_[bitbuf_ & lmask];
            int kind = (int)((e >> 13) & 7);
            if (kind == K_SUB) {
                drop((int)(e & 0xFF));
                e = lit_[(e >> 16) + (uint32_t)(bitbuf_ & ((1u << ((e >> 8) & 31)) - 1))];
                kind = (int)((e >> 13) & 7);
            }
            if (e & 0x8000u) {
                drop((int)(e & 0xFF));
                if (!text.ok[(e >> 16) & 0xFF] || (kind == K_LIT2 && !text.ok[e >> 24])) return false;
                if (bitcnt_ < 0) return false;
                continue;
            }
            if (kind == K_EOB) {
                drop((int)(e & 0xFF));
                state_ = final_block_ ? ST_TRAILER : ST_BLOCK_HEADER;
                return bitcnt_ >= 0;
            }
            if (kind != K_BASE) return false;
            drop((int)(e & 0xFF));
            drop((int)((e >> 8) & 31));
            if (bitcnt_ < 32) refill();
            uint32_t d = dist_[bitbuf_ & dmask];
            int dk = (int)((d >> 13) & 7);
            if (dk == K_SUB) {
                drop((int)(d & 0xFF));
                d = dist_[(d >> 16) + (uint32_t)(bitbuf_ & ((1u << ((d >> 8) & 31)) - 1))];
                dk = (int)((d >> 13) & 7);
            }
            if (dk != K_BASE) return false;
            drop((int)(d & 0xFF));
            drop((int)((d >> 8) & 31));
            if (bitcnt_ < 0) return false;
        }
        return false;
    }

    bool huff_block16(uint16_t *&out_ref, uint16_t *out_limit) {
        uint16_t *out = out_ref;
        const uint32_t lmask = (1u << LIT_BITS) - 1, dmask = (1u << DIST_BITS) - 1;
        while (out < out_limit) {
            refill();
            uint32_t e = lit_[bitbuf_ & lmask];
            if (e & 0x8000u) {
                drop((int)(e & 0xFF)); out[0] = (uint16_t)((e >> 16) & 0xFF); out[1] = (uint16_t)(e >> 24); out += 1 + (e >> 14 & 1);
                e = lit_[bitbuf_ & lmask];
                if (e & 0x8000u) {
                    drop((int)(e & 0xFF)); out[0] = (uint16_t)((e >> 16) & 0xFF); out[1] = (uint16_t)(e >> 24); out += 1 + (e >> 14 & 1);
                    e = lit_[bitbuf_ & lmask];
                    if (e & 0x8000u) {
                        drop((int)(e & 0xFF)); out[0] = (uint16_t)((e >> 16) & 0xFF); out[1] = (uint16_t)(e >> 24); out += 1 + (e >> 14 & 1);
                        if (bitcnt_ < 0) { out_ref = out; return fail("unexpected end of the gzip stream"); }
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
                    drop((int)(e & 0xFF)); *out++ = (uint16_t)((e >> 16) & 0xFF);
                    if (bitcnt_ < 0) { out_ref = out; return fail("unexpected end of the gzip stream"); }
                    continue;
                }
            }
            if (kind == K_EOB) {
                drop((int)(e & 0xFF));
                out_ref = out;
                if (bitcnt_ < 0) return fail("unexpected end of the gzip stream");
                state_ = final_block_ ? ST_TRAILER : ST_BLOCK_HEADER;
                return true;
            }
            if (kind != K_BASE) { out_ref = out; return fail(bitcnt_ < 15 ? "unexpected end of the gzip stream" : "invalid literal/length code"); }
            drop((int)(e & 0xFF));
            const int lx = (int)((e >> 8) & 31);
            const uint32_t length = (e >> 16) + (uint32_t)(bitbuf_ & ((1u << lx) - 1));
            drop(lx);
            if (bitcnt_ < 32) refill();
            uint32_t d = dist_[bitbuf_ & dmask];
            int dk = (int)((d >> 13) & 7);
            if (dk == K_SUB) {
                drop((int)(d & 0xFF));
                d = dist_[(d >> 16) + (uint32_t)(bitbuf_ & ((1u << ((d >> 8) & 31)) - 1))];
                dk = (int)((d >> 13) & 7);
            }
            if (dk != K_BASE) { out_ref = out; return fail(bitcnt_ < 15 ? "unexpected end of the gzip stream" : "invalid distance code"); }
            drop((int)(d & 0xFF));
            const int dx = (int)((d >> 8) & 31);
            const uint32_t distance = (d >> 16) + (uint32_t)(bitbuf_ & ((1u << dx) - 1));
            drop(dx);
            if (bitcnt_ < 0) { out_ref = out; return fail("unexpected end of the gzip stream"); }
            // distance <= 32768 = the prefix in front of every piece: the source always exists (as symbols or markers)
            const uint16_t *src = out - distance;
            uint16_t *const end = out + length;
            if (distance >= 4) {
                uint64_t w0, w1; memcpy(&w0, src, 8); memcpy(out, &w0, 8); memcpy(&w1, src + 4, 8); memcpy(out + 4, &w1, 8);
                if (length > 8) {
                    src += 8; out += 8;
                    do { uint64_t w; memcpy(&w, src, 8); memcpy(out, &w, 8); src += 4; out += 4; } while (out < end);
                }
            } else {
                do { *out++ = *src++; } while (out < end);
            }
            out = end;
        }
        out_ref = out;
        return true;
    }
};

class ParallelInflate {
public:
    static constexpr size_t WIN = MarkerInflate::WIN;

    ~ParallelInflate() { shutdown(); }

    // false: not worth it / not possible (small file, bad header) -- use FastInflate
    bool open(const uint8_t *in, size_t n, int threads, size_t segment_bytes = 1u << 20) {
        if (threads < 2 || n < 4 * segment_bytes) return false;
        MarkerInflate hdr;
        if (!hdr.skip_member_header(in, n)) return false;
        in_ = in; n_ = n; seg_bytes_ = segment_bytes;
        first_bit_ = hdr.bitpos();
        n_seg_ = (n + seg_bytes_ - 1) / seg_bytes_;
        seg_state_.reset(new std::atomic<int>[n_seg_]);
        seg_start_.assign(n_seg_, 0);
        seg_run_.assign(n_seg_, nullptr);
        for (size_t i = 0; i < n_seg_; i++) seg_state_[i].store(SEG_FREE);
        lookahead_ = (size_t)threads * 2 + 2;
        // Is this a stream the block finder can work with (dynamic blocks of text)?  Two sample segments are probed now; their
        // results are kept for the workers.  If neither has a start (binary data, stored or fixed blocks only), one thread
        // would decode everything through the 16-bit path and wait for fruitless probes on the way: slower than FastInflate.
        {
            MarkerInflate dec;
            bool any = false;
            const size_t samples[8] = {1, n_seg_ / 2, 2, n_seg_ / 2 + 1, n_seg_ / 4, 3 * n_seg_ / 4, 3, n_seg_ / 2 + 2};
            for (int i = 0; i < 8 && !(any && i >= 2); i++) {      // two samples; up to six more while nothing was found
                const size_t sseg = samples[i];
                if (sseg == 0 || sseg + 1 >= n_seg_ || preprobed_.count(sseg)) continue;
                uint64_t start = 0;
                const bool found = find_block_start(dec, sseg, &start);
                preprobed_[sseg] = found ? start : ~0ull;
                any = any || found;
            }
            if (!any) return false;
        }
        for (int t = 0; t < threads; t++) threads_.emplace_back([this] { worker(); });
        return true;
    }

    const std::string &error() const { return err_; }
    bool failed() const { return !err_.empty(); }

    long read(char *dst, size_t n) {
        size_t done = 0;
        while (done < n && !finished_ && err_.empty()) {
            if (rest_) {                                           // members after the first: serial
                const long r = rest_->read(dst + done, n - done);
                if (r < 0) { err_ = rest_->error(); break; }
                if (r == 0) { finished_ = true; break; }
                done += (size_t)r;
                continue;
            }
            if (!cur_piece_ && !next_piece()) continue;            // state changed (finished_, err_, rest_): look again
            Piece &p = *cur_piece_;
            const size_t take = std::min(n - done, p.text.size() - piece_pos_);
            memcpy(dst + done, p.text.data() + piece_pos_, take);
            piece_pos_ += take; done += take;
            if (piece_pos_ == p.text.size()) {
                crc_ = (uint32_t)crc32_combine(crc_, p.crc, (z_off_t)p.text.size());
                isize_ += (uint32_t)p.text.size();
                give_buffer(p.text);
                cur_piece_.reset();
            }
        }
        if (!err_.empty() && done == 0) return -1;
        return (long)done;
    }

private:
    enum { SEG_FREE = 0, SEG_PROBING = 1, SEG_STARTED = 2, SEG_NONE = 3 };
    struct Piece {
        std::vector<uint16_t> sym;                                 // WIN prefix symbols + the piece's symbols
        size_t n_sym = 0;
        std::vector<uint8_t> text;
        uint32_t crc = 0;
        bool resolved = false;
    };
    struct Run {                                                   // one thread's decode from one block start
        size_t seg = 0; uint64_t start_bit = 0;
        size_t settled = 0;                                        // boundary_check: segments up to here need no second look
        std::deque<std::shared_ptr<Piece>> pieces;                 // in order; guarded by mu_
        bool decoded = false;                                      // no more pieces will be added
        bool discarded = false;                                    // passed by the predecessor without a meeting: a false start
        bool confirmed = false;                                    // the predecessor arrived exactly at start_bit
        Run *next = nullptr;                                       // the run this one met (nullptr: end of member)
        bool end_of_member = false;
        uint32_t want_crc = 0, want_isize = 0; size_t next_member_offset = 0;
        std::string err;
        std::vector<uint8_t> window;                               // WIN bytes in front of start_bit (resolved), or empty for the first run
        bool window_ready = false;
    };

    const uint8_t *in_ = nullptr; size_t n_ = 0, seg_bytes_ = 0, n_seg_ = 0;
    uint64_t first_bit_ = 0;
    std::unique_ptr<std::atomic<int>[]> seg_state_;
    std::vector<uint64_t> seg_start_;
    std::vector<Run *> seg_run_;
    std::map<size_t, uint64_t> preprobed_;                         // segment -> block start found by open() (~0 = none); read-only afterwards
    std::vector<std::unique_ptr<Run>> runs_;                       // ownership; guarded by mu_
    std::mutex mu_; std::condition_variable cv_;
    std::vector<std::thread> threads_;
    size_t next_seg_ = 0;                                          // next segment to hand to a worker
    size_t consumed_seg_ = 0;                                      // segments in front of this one are not needed any more
    size_t lookahead_ = 8;
    bool stop_ = false;
    // consumer side
    Run *cur_run_ = nullptr; bool started_ = false;
    std::shared_ptr<Piece> cur_piece_; size_t piece_pos_ = 0;
    uint32_t crc_ = 0, isize_ = 0;
    bool finished_ = false;
    std::unique_ptr<FastInflate> rest_;
    std::string err_;

    void shutdown() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
        threads_.clear();
    }

    // ---- consumer: the next resolved piece in stream order; false = state changed instead
    bool next_piece() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            if (!started_) {
                cv_.wait(lk, [&] { return seg_run_[0] != nullptr || stop_; });
                if (stop_) { err_ = "closed"; return false; }
                cur_run_ = seg_run_[0]; started_ = true;
            }
            Run &r = *cur_run_;
            cv_.wait(lk, [&] { return stop_ || (!r.pieces.empty() && r.pieces.front()->resolved) || (r.pieces.empty() && r.decoded); });
            if (stop_) { err_ = "closed"; return false; }
            if (!r.pieces.empty()) {
                cur_piece_ = r.pieces.front(); r.pieces.pop_front();
                piece_pos_ = 0;
                cv_.notify_all();                                  // a long run may be waiting for room (resolve_available)
                if (cur_piece_->text.empty()) { cur_piece_.reset(); continue; }
                return true;
            }
            // the run is used up
            if (!r.err.empty()) { err_ = r.err; return false; }
            if (r.end_of_member) {
                if (crc_ != r.want_crc) { err_ = "incorrect data check"; return false; }
                if (isize_ != r.want_isize) { err_ = "incorrect length check"; return false; }
                stop_ = true;                                      // runs started inside later members are of no use
                cv_.notify_all();
                if (r.next_member_offset + 2 <= n_ && FastInflate::looks_like_gzip(in_ + r.next_member_offset, n_ - r.next_member_offset)) {
                    rest_.reset(new FastInflate());
                    rest_->reset(in_ + r.next_member_offset, n_ - r.next_member_offset);
                } else finished_ = true;                           // trailing garbage is ignored, like gzread does
                return false;
            }
            if (!r.next) { err_ = "internal: parallel inflate lost its way"; return false; }
            consumed_seg_ = r.next->seg;
            cur_run_ = r.next;
            cv_.notify_all();
        }
    }

    // ---- workers
    void worker() {
        MarkerInflate dec;
        for (;;) {
            size_t seg;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || next_seg_ >= n_seg_ || next_seg_ <= consumed_seg_ + lookahead_; });
                if (stop_ || next_seg_ >= n_seg_) return;
                seg = next_seg_++;
            }
            int expect = SEG_FREE;
            if (!seg_state_[seg].compare_exchange_strong(expect, SEG_PROBING)) continue;      // a predecessor walked over it already
            uint64_t start = 0;
            bool found = false;
            if (seg == 0) { start = first_bit_; found = true; }
            else if (preprobed_.count(seg)) { start = preprobed_.at(seg); found = start != ~0ull; }      // (const access: several workers read it)
            else found = find_block_start(dec, seg, &start);
            Run *run = nullptr;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (found) {
                    runs_.emplace_back(new Run());
                    run = runs_.back().get();
                    run->seg = seg; run->settled = seg; run->start_bit = start;
                    if (seg == 0) { run->confirmed = true; run->window_ready = true; }
                    seg_start_[seg] = start; seg_run_[seg] = run;
                    seg_state_[seg].store(SEG_STARTED);
                } else seg_state_[seg].store(SEG_NONE);
            }
            cv_.notify_all();
            if (run) decode_run(dec, *run);
        }
    }

    bool find_block_start(MarkerInflate &dec, size_t seg, uint64_t *start) {
        const uint64_t lo = (uint64_t)seg * seg_bytes_ * 8, hi = std::min<uint64_t>((uint64_t)(seg + 1) * seg_bytes_, n_ - 16) * 8;
        for (uint64_t byte = lo >> 3; byte < (hi >> 3); byte++) {
            // cheap filters on the raw bits first: BFINAL = 0, BTYPE = 2, HLIT <= 29, HDIST <= 29 (one position in nine
            // passes), then the code-length code must be complete -- Kraft sum of its HCLEN + 4 three-bit lengths exactly 1
            // (one in a few hundred of those passes); only then the real header parser and the trial decode run
            uint64_t w, w2; memcpy(&w, in_ + byte, 8); memcpy(&w2, in_ + byte + 8, 8);
            for (int s = 0; s < 8; s++) {
                const uint32_t b = (uint32_t)(w >> s);
                if ((b & 1) || ((b >> 1) & 3) != 2 || ((b >> 3) & 31) > 29 || ((b >> 8) & 31) > 29) continue;
                const int ncode = (int)((b >> 13) & 15) + 4;
                // the 3-bit lengths start 17 bits in: take them from the 128-bit window (w2:w) >> (s + 17)
                const int sh = s + 17;
                uint64_t lens = (w >> sh) | (w2 << (64 - sh));          // 64 bits: enough for 19 x 3 = 57
                uint32_t kraft = 0;
                for (int i = 0; i < ncode; i++) { const uint32_t l = (uint32_t)(lens & 7); lens >>= 3; kraft += l ? (128u >> l) : 0u; }
                if (kraft != 128u) continue;
                if (dec.probe(in_, n_, byte * 8 + (uint64_t)s)) { *start = byte * 8 + (uint64_t)s; return true; }
            }
            if ((byte & 0xFFFF) == 0) { std::lock_guard<std::mutex> lk(mu_); if (stop_) return false; }
        }
        return false;
    }

    // Buffers are recycled: a fresh multi-megabyte vector costs a zero-fill and a page fault per page, every time
    std::mutex pool_mu_;
    std::vector<std::vector<uint16_t>> sym_pool_;
    std::vector<std::vector<uint8_t>> text_pool_;
    void take_buffer(std::vector<uint16_t> &v, size_t n) {
        { std::lock_guard<std::mutex> lk(pool_mu_); if (!sym_pool_.empty()) { v.swap(sym_pool_.back()); sym_pool_.pop_back(); } }
        if (v.size() < n) v.resize(n);
    }
    void take_buffer(std::vector<uint8_t> &v, size_t n) {
        { std::lock_guard<std::mutex> lk(pool_mu_); if (!text_pool_.empty()) { v.swap(text_pool_.back()); text_pool_.pop_back(); } }
        if (v.capacity() < n) { std::vector<uint8_t>().swap(v); v.reserve(std::max(n, kPieceSyms + 600)); }
        v.resize(n);                                               // shrinking / growing inside the capacity: no reallocation
    }
    void give_buffer(std::vector<uint16_t> &v) {
        std::lock_guard<std::mutex> lk(pool_mu_);
        if (sym_pool_.size() < 64) { sym_pool_.emplace_back(); sym_pool_.back().swap(v); } else std::vector<uint16_t>().swap(v);
    }
    void give_buffer(std::vector<uint8_t> &v) {
        std::lock_guard<std::mutex> lk(pool_mu_);
        if (text_pool_.size() < 64) { text_pool_.emplace_back(); text_pool_.back().swap(v); } else std::vector<uint8_t>().swap(v);
    }

    std::shared_ptr<Piece> new_piece(const Piece *prev, bool first_of_member) {
        auto p = std::make_shared<Piece>();
        take_buffer(p->sym, WIN + kPieceSyms + 600);
        if (prev) memcpy(p->sym.data(), prev->sym.data() + prev->n_sym, WIN * 2);                 // the last WIN symbols (prefix + output are contiguous)
        else if (first_of_member) for (size_t k = 0; k < WIN; k++) p->sym[k] = MarkerInflate::NOTHING;
        else for (size_t k = 0; k < WIN; k++) p->sym[k] = (uint16_t)(256 + k);                   // markers: byte k of the unknown window
        return p;
    }
    static constexpr size_t kPieceSyms = 4u << 20;

    // what is at `bit` (a block boundary the decoder of `run` has reached)?  0 = nothing: go on; 1 = met the next run: stop
    int boundary_check(Run &run, uint64_t bit) {
        const size_t seg = (size_t)((bit >> 3) / seg_bytes_);
        // segments up to run.settled are dealt with for good (no start, or a false one): a run that walks over many
        // segments must not look at all of them again at every block
        for (size_t s = run.settled + 1; s <= seg && s < n_seg_; s++) {
            int st = seg_state_[s].load();
            if (st == SEG_FREE) {
                // Nobody has looked at this segment yet.  If the workers may take it (it is inside the look-ahead window),
                // wait for one of them to do so -- it will start a parallel run there, which is the whole point; walking
                // over it would turn this run into a serial decode of the rest of the file.  Outside the window it is mine.
                std::unique_lock<std::mutex> lk(mu_);
                // (a run its predecessor has already thrown away must not park here holding its worker: if every worker waits
                // on FREE segments nobody is left to take them)
                cv_.wait(lk, [&] { return stop_ || run.discarded || seg_state_[s].load() != SEG_FREE || s > consumed_seg_ + lookahead_; });
                if (stop_ || run.discarded) return 1;
                int expect = SEG_FREE;
                if (seg_state_[s].compare_exchange_strong(expect, SEG_NONE)) { run.settled = s; continue; }
                st = seg_state_[s].load();
            }
            if (st == SEG_PROBING) {                               // its prober may still find a start behind, at or in front of me
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || run.discarded || seg_state_[s].load() != SEG_PROBING; });
                if (stop_ || run.discarded) return 1;
                st = seg_state_[s].load();
            }
            if (st == SEG_NONE) { run.settled = s; continue; }
            // SEG_STARTED
            std::lock_guard<std::mutex> lk(mu_);
            Run *other = seg_run_[s];
            if (!other || other->discarded) { run.settled = s; continue; }
            if (other->start_bit == bit) { other->confirmed = true; run.next = other; cv_.notify_all(); return 1; }
            if (other->start_bit < bit) { other->discarded = true; run.settled = s; cv_.notify_all(); continue; }   // passed without a meeting: a false start
            return 0;                                              // it starts further on: keep going
        }
        return 0;
    }

    void decode_run(MarkerInflate &dec, Run &run) {
        dec.start_at(in_, n_, run.start_bit);
        std::shared_ptr<Piece> piece = new_piece(nullptr, run.seg == 0);
        uint16_t *out = piece->sym.data() + WIN, *limit = out + kPieceSyms;
        std::vector<std::shared_ptr<Piece>> mine;                  // this run's pieces, for the resolve step
        Resolver rs;
        auto close_piece = [&](bool last) {
            piece->n_sym = (size_t)(out - (piece->sym.data() + WIN));
            { std::lock_guard<std::mutex> lk(mu_); run.pieces.push_back(piece); if (last) run.decoded = true; }
            mine.push_back(piece);
            cv_.notify_all();
        };
        bool first_block = true;
        for (;;) {
            if (dec.at_block_header() && !first_block) {
                bool dropped;
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    dropped = stop_ || run.discarded;
                    if (dropped) run.err = "discarded";
                }
                if (dropped) {                                           // a false start (or the end): its pieces are of no use
                    std::lock_guard<std::mutex> lk(mu_);
                    run.pieces.clear(); run.decoded = true;
                    return;
                }
                const uint64_t bit = dec.bitpos();
                if ((size_t)((bit >> 3) / seg_bytes_) > run.seg && boundary_check(run, bit)) { close_piece(true); break; }
            }
            first_block = false;
            if (out >= limit) {                                     // piece full: continue in a new one that starts with my last WIN symbols
                close_piece(false);
                auto nxt = new_piece(nullptr, false);
                memcpy(nxt->sym.data(), out - WIN, WIN * 2);
                piece = nxt;
                resolve_available(rs, run, mine);                 // (gives the closed piece's symbols back: copy the prefix first)
                out = piece->sym.data() + WIN; limit = out + kPieceSyms;
            }
            if (!dec.decode_some(out, limit)) {
                std::lock_guard<std::mutex> lk(mu_);
                run.err = dec.error();
                piece->n_sym = (size_t)(out - (piece->sym.data() + WIN));
                run.pieces.push_back(piece); mine.push_back(piece); run.decoded = true;
                break;
            }
            if (dec.at_final_block_end()) {
                uint32_t c = 0, isz = 0; size_t nxt = 0;
                const bool ok = dec.read_trailer(&c, &isz, &nxt);
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    if (ok) { run.end_of_member = true; run.want_crc = c; run.want_isize = isz; run.next_member_offset = nxt; }
                    else run.err = "unexpected end of the gzip stream";
                }
                close_piece(true);
                break;
            }
        }
        cv_.notify_all();
        resolve_run(rs, run, mine);
    }

    // symbol -> byte once the window in front of a run is known: 0..255 themselves, 256 + k = byte k of that window.
    // NOTHING, and every marker in the member's first run (it has no window), is a copy from in front of the member's
    // first byte: an error.
    struct Resolver {
        std::vector<uint8_t> lut;
        uint32_t bias = 0;                                         // v + bias >= 0x10000  <=>  v is not resolvable
        std::vector<uint8_t> tail;                                 // running "last WIN bytes of the run's text so far"
        bool ready = false;
        size_t done = 0;                                           // pieces of the run resolved so far
    };
    void prepare(Resolver &rs, const Run &run) {                   // run.window_ready holds
        rs.lut.assign(65536, 0);
        for (int v = 0; v < 256; v++) rs.lut[(size_t)v] = (uint8_t)v;
        const bool have_window = !run.window.empty();
        if (have_window) memcpy(rs.lut.data() + 256, run.window.data(), WIN);
        rs.bias = have_window ? 0x10000u - (256u + (uint32_t)WIN) : 0x10000u - 256u;
        rs.tail = run.window;
        rs.ready = true;
    }
    void resolve_piece(Resolver &rs, Run &run, Piece &p) {
        take_buffer(p.text, p.n_sym);
        const uint16_t *s = p.sym.data() + WIN;
        uint8_t *t = p.text.data();
        const uint8_t *lut = rs.lut.data();
        const uint32_t bias = rs.bias;
        uint32_t acc = 0;
        size_t j = 0;
        for (; j + 4 <= p.n_sym; j += 4) {
            const uint32_t a = s[j], b = s[j + 1], c = s[j + 2], d = s[j + 3];
            t[j] = lut[a]; t[j + 1] = lut[b]; t[j + 2] = lut[c]; t[j + 3] = lut[d];
            acc |= (a + bias) | (b + bias) | (c + bias) | (d + bias);
        }
        for (; j < p.n_sym; j++) { t[j] = lut[s[j]]; acc |= s[j] + bias; }
        const bool bad = (acc & 0x10000u) != 0;
        give_buffer(p.sym);
        p.crc = crc32_update(0, p.text.data(), p.text.size());
        if (p.text.size() >= WIN) rs.tail.assign(p.text.end() - WIN, p.text.end());
        else { rs.tail.insert(rs.tail.end(), p.text.begin(), p.text.end()); if (rs.tail.size() > WIN) rs.tail.erase(rs.tail.begin(), rs.tail.end() - WIN); }
        std::lock_guard<std::mutex> lk(mu_);
        if (bad && run.err.empty()) { run.err = "invalid distance too far back"; p.text.clear(); }
        p.resolved = true;
        cv_.notify_all();
    }
    // during the decode: resolve what can be resolved already (a long run must not pile up 16-bit pieces), and do not run
    // away from a slow consumer
    void resolve_available(Resolver &rs, Run &run, std::vector<std::shared_ptr<Piece>> &mine) {
        if (!rs.ready) {
            std::lock_guard<std::mutex> lk(mu_);
            if (!run.window_ready) return;
        }
        if (!rs.ready) prepare(rs, run);
        for (; rs.done < mine.size(); rs.done++) resolve_piece(rs, run, *mine[rs.done]);
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || run.discarded || run.pieces.size() < 8; });
    }

    // after the decode: the rest of the pieces, once the window in front of the run is known; publishes the next window
    void resolve_run(Resolver &rs, Run &run, std::vector<std::shared_ptr<Piece>> &mine) {
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || run.discarded || run.window_ready; });
            if (stop_ || run.discarded) { run.pieces.clear(); return; }
        }
        if (!rs.ready) prepare(rs, run);
        // The next run is waiting for its window = my last WIN bytes: resolve those first and hand them over, so that the
        // runs wait for 32 KiB of their predecessor, not for all of it
        bool next_served = false;
        if (run.next && !mine.empty() && mine.back()->n_sym >= WIN && rs.done < mine.size()) {
            const Piece &lp = *mine.back();
            const uint16_t *s = lp.sym.data() + WIN + (lp.n_sym - WIN);
            std::vector<uint8_t> w(WIN);
            for (size_t j = 0; j < WIN; j++) w[j] = rs.lut[s[j]];
            std::lock_guard<std::mutex> lk(mu_);
            run.next->window.swap(w);
            run.next->window_ready = true;
            next_served = true;
            cv_.notify_all();
        }
        for (; rs.done < mine.size(); rs.done++) resolve_piece(rs, run, *mine[rs.done]);
        std::lock_guard<std::mutex> lk(mu_);
        if (run.next && !next_served) {
            // a window shorter than WIN (text so far < 32 KiB): right-align it, what lies in front does not exist
            run.next->window.assign(WIN, 0);
            memcpy(run.next->window.data() + (WIN - rs.tail.size()), rs.tail.data(), rs.tail.size());
            run.next->window_ready = true;
        }
        std::vector<uint8_t>().swap(run.window);                   // (a 100-GB file has 10^5 runs: do not keep 32 KiB for each)
        cv_.notify_all();
    }
};

// BGZF (bgzip, htslib; RFC 1952 members of at most 64 KiB whose extra field 'BC' holds the member's size): the member
// boundaries are known without decoding, so groups of members go to the threads as they are -- each group is a small
// multi-member gzip file for FastInflate (which checks every member's CRC-32 and ISIZE) -- and come back in order.
class BgzfInflate {
public:
    ~BgzfInflate() { shutdown(); }

    // size of the BGZF member at p (0 = not a BGZF member)
    static size_t member_size(const uint8_t *p, size_t left) {
        if (left < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
        const size_t xlen = p[10] | (size_t)p[11] << 8;
        if (left < 12 + xlen) return 0;
        for (size_t q = 12; q + 4 <= 12 + xlen;) {
            const size_t slen = p[q + 2] | (size_t)p[q + 3] << 8;
            if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen) {
                const size_t bsize = (p[q + 4] | (size_t)p[q + 5] << 8) + 1;
                return bsize >= 12 + xlen + 8 && bsize <= left ? bsize : 0;
            }
            q += 4 + slen;
        }
        return 0;
    }

    bool open(const uint8_t *in, size_t n, int threads, size_t group_bytes = 4u << 20) {
        if (threads < 2 || n < 2 * group_bytes || !member_size(in, n)) return false;
        in_ = in; n_ = n; group_bytes_ = group_bytes;
        max_ahead_ = (size_t)threads * 2 + 2;
        for (int t = 0; t < threads; t++) threads_.emplace_back([this] { worker(); });
        return true;
    }
    const std::string &error() const { return err_; }
    bool failed() const { return !err_.empty(); }

    long read(char *dst, size_t n) {
        size_t done = 0;
        while (done < n && err_.empty() && !finished_) {
            if (rest_) {                                           // something that is not BGZF follows: serial
                const long r = rest_->read(dst + done, n - done);
                if (r < 0) { err_ = rest_->error(); break; }
                if (r == 0) { finished_ = true; break; }
                done += (size_t)r;
                continue;
            }
            if (!cur_) {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return (!groups_.empty() && groups_.front()->ready) || (groups_.empty() && walked_all_); });
                if (groups_.empty()) {
                    if (walk_pos_ < n_ && FastInflate::looks_like_gzip(in_ + walk_pos_, n_ - walk_pos_)) {
                        rest_.reset(new FastInflate()); rest_->reset(in_ + walk_pos_, n_ - walk_pos_);
                    } else finished_ = true;                       // end of file, or trailing garbage (ignored, like gzread)
                    continue;
                }
                cur_ = groups_.front(); groups_.pop_front(); pos_ = 0;
                consumed_++;
                cv_.notify_all();
                if (!cur_->err.empty()) { err_ = cur_->err; break; }
            }
            const size_t take = std::min(n - done, cur_->text.size() - pos_);
            memcpy(dst + done, cur_->text.data() + pos_, take);
            pos_ += take; done += take;
            if (pos_ == cur_->text.size()) cur_.reset();
        }
        if (!err_.empty() && done == 0) return -1;
        return (long)done;
    }

private:
    struct Group { size_t begin = 0, end = 0; std::vector<char> text; std::string err; bool ready = false; };
    const uint8_t *in_ = nullptr; size_t n_ = 0, group_bytes_ = 0;
    std::mutex mu_; std::condition_variable cv_;
    std::vector<std::thread> threads_;
    std::deque<std::shared_ptr<Group>> groups_;                    // in file order
    size_t walk_pos_ = 0, handed_ = 0, consumed_ = 0, max_ahead_ = 8;
    bool walked_all_ = false, stop_ = false;
    std::shared_ptr<Group> cur_; size_t pos_ = 0;
    std::unique_ptr<FastInflate> rest_;
    bool finished_ = false;
    std::string err_;

    void shutdown() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
        threads_.clear();
    }
    void worker() {
        FastInflate fi;
        for (;;) {
            std::shared_ptr<Group> g;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || walked_all_ || handed_ < consumed_ + max_ahead_; });
                if (stop_ || walked_all_) return;
                // the next group: whole members from walk_pos_ on, about group_bytes_ of them
                size_t p = walk_pos_;
                while (p < n_ && p - walk_pos_ < group_bytes_) {
                    const size_t sz = member_size(in_ + p, n_ - p);
                    if (!sz) break;
                    p += sz;
                }
                if (p == walk_pos_) { walked_all_ = true; cv_.notify_all(); return; }      // end of the BGZF part
                g = std::make_shared<Group>();
                g->begin = walk_pos_; g->end = p;
                walk_pos_ = p; handed_++;
                groups_.push_back(g);
            }
            fi.reset(in_ + g->begin, g->end - g->begin);
            std::vector<char> text;
            text.resize((g->end - g->begin) * 5 + (1u << 16));
            size_t have = 0;
            std::string err;
            for (;;) {
                if (text.size() - have < (1u << 20)) text.resize(text.size() * 2);
                const long r = fi.read(text.data() + have, text.size() - have);
                if (r < 0) { err = fi.error(); break; }
                if (r == 0) break;
                have += (size_t)r;
            }
            text.resize(have);
            std::lock_guard<std::mutex> lk(mu_);
            g->text.swap(text); g->err = err; g->ready = true;
            cv_.notify_all();
        }
    }
};

}  // namespace mfkc
