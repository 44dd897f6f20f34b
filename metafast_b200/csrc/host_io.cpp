// host_io.cpp -- CPU side of the path behind the same C ABI (no GPU needed):
//   * FASTA / FASTQ (+ .gz) readers with the reference's parser rules
//       [itmo]/io/ReadersUtils.java:27-54,57-77,85-102
//       [itmo]/io/readers/FastaReader.java:54-108, FastqReader.java:53-114,
//       [itmo]/io/readers/FastaReaderFromXQSource.java:62-85,
//       [itmo]/io/formats/Illumina.java:7-12, Sanger.java:7-12, [itmo]/dna/DnaQ.java:140-150
//   * .stat.txt writer ([itmo]/statistics/QuickQuantitativeStatistics.java:38-72)
//   * the synthetic read generator's host half (synth.h)
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mfkc.h"
#include "fast_inflate.h"
#include "parallel_inflate.h"
#include "synth.h"

namespace {

bool ends_with_ci(const std::string &s, const char *suf) {
    const size_t n = s.size(), m = strlen(suf);
    return n >= m && strcasecmp(s.c_str() + n - m, suf) == 0;
}
std::string base_name(const std::string &p) {
    const size_t i = p.find_last_of('/');
    return i == std::string::npos ? p : p.substr(i + 1);
}
// [itmo]/utils/FileUtils.java:199-210: strip the first matching extension (case-insensitive)
std::string remove_extension(const std::string &s, std::initializer_list<const char *> exts) {
    for (const char *e : exts)
        if (ends_with_ci(s, e)) return s.substr(0, s.size() - strlen(e));
    return s;
}

enum Format { F_UNKNOWN = 0, F_FASTA, F_FASTA_GZ, F_FASTQ, F_FASTQ_GZ, F_OTHER };

// [itmo]/io/ReadersUtils.java:27-54
Format detect_format(const std::string &path) {
    std::string name = base_name(path);
    bool gz = false;
    if (ends_with_ci(name, ".gz")) { gz = true; name = name.substr(0, name.size() - 3); }
    if (ends_with_ci(name, ".bz2")) return F_OTHER;                 // out of scope: delegate to the library
    if (ends_with_ci(name, ".binq")) return F_OTHER;
    if (ends_with_ci(name, ".fastq") || ends_with_ci(name, ".fq")) return gz ? F_FASTQ_GZ : F_FASTQ;
    if (ends_with_ci(name, ".fasta") || ends_with_ci(name, ".fa") || ends_with_ci(name, ".fn") || ends_with_ci(name, ".fna"))
        return gz ? F_FASTA_GZ : F_FASTA;
    return F_UNKNOWN;
}

// NamedSource.name(): FastaReader.java:22, FastaGZReader.java:17, FastqReader.java:25, FastqGZReader.java:21
std::string library_name(const std::string &path, Format f) {
    const std::string b = base_name(path);
    switch (f) {
        case F_FASTA: return remove_extension(b, {".fasta", ".fa", ".fn", ".fna"});
        case F_FASTA_GZ: return remove_extension(b, {".fasta.gz", ".fa.gz", ".fn.gz", ".fna.gz"});
        case F_FASTQ: return remove_extension(b, {".fastq", ".fq"});
        case F_FASTQ_GZ: return remove_extension(b, {".fastq.gz", ".fq.gz"});
        default: return b;
    }
}

// first '\n' or '\r' in [q, end), or end
static inline const char *find_eol(const char *q, const char *end) {
#if defined(__x86_64__)
    const __m128i nl = _mm_set1_epi8('\n'), cr = _mm_set1_epi8('\r');
    while (q + 16 <= end) {
        const __m128i v = _mm_loadu_si128((const __m128i *)q);
        const unsigned m = (unsigned)_mm_movemask_epi8(_mm_or_si128(_mm_cmpeq_epi8(v, nl), _mm_cmpeq_epi8(v, cr)));
        if (m) return q + __builtin_ctz(m);
        q += 16;
    }
#endif
    while (q < end && *q != '\n' && *q != '\r') q++;
    return q;
}

// FASTQ record fast path: true when every base is one of AaCcGgTt and every quality char q has lo < q <= 126 and
// q != lo + 64, i.e. the per-character rules of RecordParser::fastq_record raise no error and keep the read (phred != 0
// everywhere, no N).  Anything else -- N, '.', phred 0, bad characters -- is left to the exact scalar loop.
static inline bool fastq_record_is_clean(const char *d, const char *q, size_t n, int lo) {
    size_t i = 0;
#if defined(__x86_64__)
    const __m128i up = _mm_set1_epi8((char)0xDF), A = _mm_set1_epi8('A'), Cc = _mm_set1_epi8('C'), G = _mm_set1_epi8('G'), T = _mm_set1_epi8('T');
    const __m128i vlo = _mm_set1_epi8((char)lo), v127 = _mm_set1_epi8(127), vz = _mm_set1_epi8((char)(lo + 64));
    for (; i + 16 <= n; i += 16) {
        const __m128i x = _mm_and_si128(_mm_loadu_si128((const __m128i *)(d + i)), up);
        const __m128i okb = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(x, A), _mm_cmpeq_epi8(x, Cc)), _mm_or_si128(_mm_cmpeq_epi8(x, G), _mm_cmpeq_epi8(x, T)));
        const __m128i v = _mm_loadu_si128((const __m128i *)(q + i));
        // signed compares: bytes >= 128 are negative and fail v > lo
        const __m128i okq = _mm_andnot_si128(_mm_cmpeq_epi8(v, vz), _mm_and_si128(_mm_cmpgt_epi8(v, vlo), _mm_cmpgt_epi8(v127, v)));
        if (_mm_movemask_epi8(_mm_and_si128(okb, okq)) != 0xFFFF) return false;
    }
#endif
    for (; i < n; i++) {
        const unsigned char b = (unsigned char)d[i] & 0xDFu, c = (unsigned char)q[i];
        if (!(b == 'A' || b == 'C' || b == 'G' || b == 'T')) return false;
        if (!(c > lo && c <= 126 && c != lo + 64)) return false;
    }
    return true;
}

// BufferedReader.readLine over a (possibly gzip) stream: terminators \n, \r\n, \r.
class LineReader {
public:
    static constexpr bool kStableLines = false;      // a line is valid until the next call only
    bool open(const std::string &path) {
        close();
        f_ = gzopen(path.c_str(), "rb");           // transparent for plain files (GZIPInputStream otherwise)
        if (!f_) return false;
        gzbuffer(f_, 1 << 20);
        buf_.resize(1 << 22);
        pos_ = len_ = 0; eof_ = false; pending_cr_ = false;
        return true;
    }
    void close() { if (f_) { gzclose(f_); f_ = nullptr; } }
    ~LineReader() { close(); }
    bool failed() const { return io_error_; }

    // Returns false at end of stream.  `line` stays valid until the next call.
    bool next(const char *&line, size_t &len) {
        acc_.clear();
        bool have_acc = false;
        for (;;) {
            if (pos_ == len_) {
                if (!fill()) {
                    if (have_acc) { line = acc_.data(); len = acc_.size(); return true; }
                    return false;
                }
            }
            if (pending_cr_) {                      // a lone \r ended the previous line; swallow a following \n
                pending_cr_ = false;
                if (buf_[pos_] == '\n') { pos_++; continue; }
            }
            const char *p = buf_.data() + pos_;
            const size_t avail = len_ - pos_;
            const size_t i = (size_t)(find_eol(p, p + avail) - p);
            if (i < avail) {
                const bool cr = p[i] == '\r';
                if (have_acc) { acc_.append(p, i); line = acc_.data(); len = acc_.size(); }
                else { line = p; len = i; }
                pos_ += i + 1;
                if (cr) pending_cr_ = true;
                return true;
            }
            acc_.append(p, avail);                  // line continues in the next buffer
            have_acc = true;
            pos_ = len_;
        }
    }

private:
    bool fill() {
        if (eof_) return false;
        const int r = gzread(f_, buf_.data(), (unsigned)buf_.size());
        if (r < 0) { io_error_ = true; eof_ = true; return false; }
        if (r == 0) { eof_ = true; return false; }
        pos_ = 0; len_ = (size_t)r;
        return true;
    }
    gzFile f_ = nullptr;
    std::vector<char> buf_;
    size_t pos_ = 0, len_ = 0;
    bool eof_ = false, pending_cr_ = false, io_error_ = false;
    std::string acc_;
};

struct CodeTable {
    bool ok[256];
    CodeTable() {
        memset(ok, 0, sizeof ok);
        for (const char *p = "AaCcGgTt"; *p; p++) ok[(unsigned char)*p] = true;   // [itmo]/dna/DnaTools.java:46-60
    }
};
const CodeTable kCode;

}  // namespace

// The same line rules over a block of text that is already in memory (a chunk of whole records).
class MemLineReader {
public:
    static constexpr bool kStableLines = true;       // lines point into the block
    MemLineReader(const char *p, size_t n) : p_(p), end_(p + n) {}
    bool failed() const { return false; }
    bool next(const char *&line, size_t &len) {
        if (p_ == end_) return false;
        const char *q = find_eol(p_, end_);
        line = p_; len = (size_t)(q - p_);
        if (q == end_) { p_ = end_; return true; }               // last line without a terminator
        p_ = q + 1;
        if (*q == '\r' && p_ < end_ && *p_ == '\n') p_++;        // \r\n
        return true;
    }
private:
    const char *p_, *end_;
};

// Record-level rules of the reference parsers over any line source.
template <class LR>
struct RecordParser {
    Format fmt = F_UNKNOWN;
    int phred_lo = 64;                     // Illumina (64) unless the sniff says Sanger (33)
    uint64_t all_reads = 0, skipped = 0;
    std::string err;
    std::string sb;                        // FASTA record assembly
    bool fasta_eof = false;

    // ---- FASTA: FastaReader.java:82-104 (readNextDataLine) + :54-66 (N-drop)
    // returns 1 = record in `out`, 0 = end, <0 error
    int fasta_next(LR &lr, std::string &out) {
        for (;;) {
            if (fasta_eof) return 0;
            sb.clear();
            for (;;) {
                const char *l; size_t n;
                if (!lr.next(l, n)) { fasta_eof = true; break; }
                if (n > 0 && (l[0] == '>' || l[0] == ';')) { if (!sb.empty()) break; }
                else sb.append(l, n);
            }
            if (lr.failed()) { err = "read error (corrupt gzip stream?)"; return MFKC_E_IO; }
            if (sb.empty()) continue;
            all_reads++;
            if (sb.find('N') != std::string::npos || sb.find('n') != std::string::npos) { skipped++; continue; }
            for (unsigned char c : sb)
                if (!kCode.ok[c]) {           // DnaTools.fromChar throws IllegalArgumentException (IUPAC codes are
                    err = std::string("Incorrect nucleotide char: \"") + (char)c + "\"";   // randomised there: unsupported)
                    return MFKC_E_FORMAT;
                }
            out.swap(sb);
            return 1;
        }
    }

    // ---- FASTQ: FastqReader.java:84-110; 1 = line, 0 = EOF, <0 error
    int fastq_data_line(LR &r, const char *&l, size_t &n) {
        bool ok = r.next(l, n);
        while (ok && n == 0) ok = r.next(l, n);           // skipping empty lines
        if (!ok) return 0;
        if (!(l[0] == '@' || l[0] == '+')) {
            err = "Unknown structure of fastq file! Waiting \"@ID\" or \"+ID\" string, found \"" +
                  std::string(l, n > 20 ? 20 : n) + (n > 20 ? "..." : "") + "\".";
            return MFKC_E_FORMAT;
        }
        if (!r.next(l, n)) { err = "Unexpected end of file. File is corrupted/Format mismatch."; return MFKC_E_FORMAT; }
        return 1;
    }

    // One FASTQ record (FastqReader.java:53-82 + FastaReaderFromXQSource.java:62-69).
    // returns 1 = record parsed (kept says whether it survives), 0 = end, <0 error (-100 = illegal quality)
    int fastq_record(LR &r, int lo, std::string *out, bool &kept, std::vector<uint8_t> *append = nullptr) {
        const char *l; size_t n;
        int s = fastq_data_line(r, l, n);
        if (s <= 0) return s;
        const char *d = l; const size_t dn = n;
        if (!LR::kStableLines) { data_.assign(l, n); d = data_.data(); }
        s = fastq_data_line(r, l, n);
        if (s == 0) { err = "Unexpected end of file. File is corrupted/Format mismatch."; return MFKC_E_FORMAT; }
        if (s < 0) return s;
        if (n != dn) {
            err = "Bad DnaQ record: length of chars and quality is not the same. File is corrupted/Format mismatch.";
            return MFKC_E_FORMAT;
        }
        kept = true;
        if (!fastq_record_is_clean(d, l, n, lo))
            for (size_t i = 0; i < n; i++) {
                const unsigned char ch = (unsigned char)d[i];
                if (ch == 'N' || ch == 'n' || ch == '.') { kept = false; continue; }    // appendUnknown: phred 0
                if (!kCode.ok[ch]) { err = std::string("Incorrect nucleotide char: \"") + (char)ch + "\""; return MFKC_E_FORMAT; }
                const int q = (unsigned char)l[i];
                if (q < lo || q > 126) {                                                // Illumina.java:8 / Sanger.java:8
                    err = std::string("Invalid quality code char: \"") + (char)q + "\"";
                    return -100;
                }
                if (((q - lo) & 63) == 0) kept = false;       // 6-bit phred (DnaQ.java:140-150) == 0 -> read dropped
            }
        if (kept) {
            if (append) append->insert(append->end(), (const uint8_t *)d, (const uint8_t *)d + dn);
            else if (out) out->assign(d, dn);
        }
        return 1;
    }

    // next kept read into cur; 1 / 0 / <0
    // `append`: the kept read goes to the end of that vector instead of `cur` (the chunk workers: no per-read string)
    int next_read(LR &lr, std::string &cur, std::vector<uint8_t> *append = nullptr) {
        int s;
        if (fmt == F_FASTA || fmt == F_FASTA_GZ) {
            s = fasta_next(lr, cur);
            if (s > 0 && append) append->insert(append->end(), cur.begin(), cur.end());
            return s;
        }
        for (;;) {
            bool kept = false;
            s = fastq_record(lr, phred_lo, &cur, kept, append);
            if (s == -100) s = MFKC_E_FORMAT;
            if (s <= 0) break;
            all_reads++;
            if (kept) return 1;
            skipped++;
        }
        if (s == 0 && lr.failed()) { err = "read error (corrupt gzip stream?)"; s = MFKC_E_IO; }
        return s;
    }

private:
    std::string data_;
};

// ------------------------------------------------------------------------------------------
// Parallel ingest.  The reference parses inside the synchronized dispatcher: one thread
// gunzips, splits and validates while the workers wait (src/io/ReadsDispatcher.java:34-38).  Here one producer
// thread reads / inflates the file and cuts the text into chunks of WHOLE records with a line-level state machine
// (FASTQ: [empty lines] header, data, [empty lines] '+' line, quality; FASTA: cut in front of a header line);
// P worker threads run the unchanged record rules (RecordParser) on their chunks; mfkc_reader_next hands the
// chunks out in file order.  Results -- kept reads, their order, counters, the first error -- are those of
// the serial parser; MFKC_READER_THREADS=1 selects the serial path.
// ------------------------------------------------------------------------------------------
struct IngestChunk {
    std::vector<char> text;                                     // owned text buffer (inflated / read data); its size is a capacity
    const char *tptr = nullptr; size_t tlen = 0;                // the chunk's text: inside `text`, or a view of the mapped file
    std::vector<uint8_t> bases; std::vector<uint64_t> offs;     // offs[0] = 0
    uint64_t all = 0, skipped = 0;
    int status = 0; std::string err;                            // status < 0: the error follows the reads of this chunk
    bool parsed = false, last = false;
};

// FASTQ fast path of last_boundary.  A block that starts at a record boundary and holds nothing but non-empty lines ending
// in '\n' (no '\r' anywhere, no empty line) leaves the state machine below in state (line number mod 4), so the last
// boundary follows from the NUMBER of lines: one vectorised pass instead of two memchr calls per line.  Returns false when
// the block is not of that shape (CRLF files, blank lines, ...): the exact machine then decides.
static bool fastq_boundary_by_line_count(const char *p, size_t n, size_t *cut) {
    if (!n || p[0] == '\n') return false;
    size_t lines = 0, i = 1;                                    // byte 0 is not a newline: every position below has a left neighbour
    if (p[0] == '\r') return false;
#if defined(__x86_64__)
    // Plain SSE2, no movemask / popcount per vector (the baseline x86-64 target has no popcnt instruction: the builtin is
    // a dozen ALU operations): newlines are counted in 16 byte-wide counters (compare result = -1 per hit), folded with
    // psadbw every 255 vectors; "\n\n" = a newline whose left neighbour is one (second load, shifted by one byte);
    // '\r' hits are ORed together.  Both are looked at when the counters are folded.
    const __m128i nl = _mm_set1_epi8('\n'), cr = _mm_set1_epi8('\r'), zero = _mm_setzero_si128();
    __m128i bad = zero;
    while (i + 16 <= n) {
        __m128i acc = zero;
        const size_t stop = std::min(n - 15, i + 255 * 16);      // i + 16 <= n  <=>  i < n - 15
        for (; i < stop; i += 16) {
            const __m128i v = _mm_loadu_si128((const __m128i *)(p + i));
            const __m128i left = _mm_loadu_si128((const __m128i *)(p + i - 1));
            const __m128i is_nl = _mm_cmpeq_epi8(v, nl);
            bad = _mm_or_si128(bad, _mm_or_si128(_mm_cmpeq_epi8(v, cr), _mm_and_si128(is_nl, _mm_cmpeq_epi8(left, nl))));
            acc = _mm_sub_epi8(acc, is_nl);
        }
        const __m128i sums = _mm_sad_epu8(acc, zero);
        lines += (size_t)_mm_cvtsi128_si64(sums) + (size_t)_mm_cvtsi128_si64(_mm_unpackhi_epi64(sums, sums));
        if (_mm_movemask_epi8(bad)) return false;
    }
#endif
    for (; i < n; i++) {
        const char c = p[i];
        if (c == '\r') return false;
        if (c == '\n') { if (p[i - 1] == '\n') return false; lines++; }
    }
    if (lines < 4) { *cut = 0; return true; }
    // the cut is just behind the newline that ends line 4 * (lines / 4): step back over the surplus newlines
    const char *q = (const char *)memrchr(p, '\n', n);
    for (size_t r = lines & 3; r; r--) q = (const char *)memrchr(p, '\n', (size_t)(q - p));
    *cut = (size_t)(q - p) + 1;
    return true;
}

// Scout's view of one block [lo, hi) of a text that starts at `base`: number of '\n' in it, or -1 when the block holds a
// '\r' or an empty line ("\n\n", also across its left edge; a '\n' as the very first byte of the text) -- then the exact
// state machine has to decide.  Same vector body as above.
static int64_t count_clean_newlines(const char *base, size_t lo, size_t hi) {
    const char *p = base;
    size_t i = lo, lines = 0;
    if (i == 0) {
        if (hi == 0) return 0;
        if (p[0] == '\n' || p[0] == '\r') return -1;
        i = 1;
    }
#if defined(__x86_64__)
    const __m128i nl = _mm_set1_epi8('\n'), cr = _mm_set1_epi8('\r'), zero = _mm_setzero_si128();
    __m128i bad = zero;
    while (i + 16 <= hi) {
        __m128i acc = zero;
        const size_t stop = std::min(hi - 15, i + 255 * 16);
        for (; i < stop; i += 16) {
            const __m128i v = _mm_loadu_si128((const __m128i *)(p + i));
            const __m128i left = _mm_loadu_si128((const __m128i *)(p + i - 1));
            const __m128i is_nl = _mm_cmpeq_epi8(v, nl);
            bad = _mm_or_si128(bad, _mm_or_si128(_mm_cmpeq_epi8(v, cr), _mm_and_si128(is_nl, _mm_cmpeq_epi8(left, nl))));
            acc = _mm_sub_epi8(acc, is_nl);
        }
        const __m128i sums = _mm_sad_epu8(acc, zero);
        lines += (size_t)_mm_cvtsi128_si64(sums) + (size_t)_mm_cvtsi128_si64(_mm_unpackhi_epi64(sums, sums));
        if (_mm_movemask_epi8(bad)) return -1;
    }
#endif
    for (; i < hi; i++) {
        const char c = p[i];
        if (c == '\r') return -1;
        if (c == '\n') { if (p[i - 1] == '\n') return -1; lines++; }
    }
    return (int64_t)lines;
}

// offset of the last record boundary in [p, p+n) (0 = none); `eof`: the text ends here
static size_t last_boundary(const char *p, size_t n, bool fastq, bool eof) {
    size_t fast_cut = 0;
    if (fastq && !eof && fastq_boundary_by_line_count(p, n, &fast_cut)) return fast_cut;
    size_t pos = 0, cut = 0;
    int state = 0;                                              // FASTQ: 0 header, 1 data, 2 '+' line, 3 quality
    while (pos < n) {
        // line terminators: \n, \r\n and a lone \r (BufferedReader.readLine); memchr does the scanning
        const char *nl = (const char *)memchr(p + pos, '\n', n - pos);
        const size_t lim = nl ? (size_t)(nl - p) : n;
        const char *cr = (const char *)memchr(p + pos, '\r', lim - pos);
        size_t q, next;
        if (cr) {
            q = (size_t)(cr - p); next = q + 1;
            if (next == n && !eof) break;                       // a '\n' may still follow
            if (next < n && p[next] == '\n') next++;
        } else {
            if (!nl) break;                                     // unterminated line: belongs to the next chunk
            q = lim; next = q + 1;
        }
        const size_t len = q - pos;
        if (fastq) {
            if (state == 0) { if (len) state = 1; }
            else if (state == 1) state = 2;
            else if (state == 2) { if (len) state = 3; }
            else { state = 0; cut = next; }
        } else if (len && (p[pos] == '>' || p[pos] == ';') && pos) cut = pos;
        pos = next;
    }
    return cut;
}

// FASTA flavour with a resume point: the last line start in [max(from, 1), n) that opens a header ('>' or ';'), 0 = none.
// Positions below `from` are known to hold none (an earlier call on a shorter prefix returned 0), so a record of L bytes
// costs O(L) to step over while the block is widened chunk by chunk, not O(L^2 / chunk).
static size_t fasta_boundary_from(const char *p, size_t from, size_t n) {
    if (from < 1) from = 1;
    size_t end = n;
    while (end > from) {
        const char *a = (const char *)memrchr(p + from, '>', end - from);
        const char *b = (const char *)memrchr(p + from, ';', end - from);
        const char *q = !a ? b : !b ? a : (a > b ? a : b);
        if (!q) return 0;
        const size_t i = (size_t)(q - p);
        if (p[i - 1] == '\n' || p[i - 1] == '\r') return i;
        end = i;
    }
    return 0;
}

struct mfkc_reader {
    std::string path, name, err;
    Format fmt = F_UNKNOWN;
    int phred_lo = 64;
    uint64_t all_reads = 0, skipped = 0;
    bool done = false;
    // serial path
    LineReader lr;
    RecordParser<LineReader> sp;
    std::string cur; bool have_cur = false;
    // parallel path
    int n_threads = 1;
    std::vector<std::thread> threads;
    std::mutex mu; std::condition_variable cv_work, cv_done, cv_space;
    std::deque<std::shared_ptr<IngestChunk>> ordered;            // chunks in file order (consumer pops the front)
    std::deque<std::shared_ptr<IngestChunk>> todo;               // not parsed yet
    bool producer_done = false, stop = false;
    std::shared_ptr<IngestChunk> cur_chunk; size_t cur_read = 0;
    size_t kChunkText = 4u << 20;                               // MFKC_READER_CHUNK (bytes): small chunks for the tests
    static constexpr size_t kMaxQueued = 48;

    bool is_fastq() const { return fmt == F_FASTQ || fmt == F_FASTQ_GZ; }

    // ReadersUtils.determineQualityFormat :63-77
    int sniff_quality() {
        LineReader r;
        if (!r.open(path)) { err = "can't open " + path; return MFKC_E_IO; }
        RecordParser<LineReader> p; p.fmt = fmt;
        for (int i = 0; i < 1000; i++) {
            bool kept;
            const int s = p.fastq_record(r, 64, nullptr, kept);
            if (s == 0) break;
            if (s == -100) { phred_lo = 33; return MFKC_OK; }
            if (s < 0) { err = p.err; return s; }
        }
        phred_lo = 64;
        return MFKC_OK;
    }

    // serial: next kept read into cur; 1 / 0 / <0
    int advance() {
        if (done) return 0;
        const int s = sp.next_read(lr, cur);
        all_reads = sp.all_reads; skipped = sp.skipped;
        if (s < 0) err = sp.err;
        if (s <= 0) done = true;
        return s;
    }

    // ---- parallel path
    // Chunk objects are recycled (their vectors keep their capacity): a fresh 4-MB vector per chunk costs a page fault
    // per page and an munmap when the worker drops it, which was a third of the producer's time.
    std::vector<std::shared_ptr<IngestChunk>> pool;
    std::shared_ptr<IngestChunk> fresh_chunk() {
        std::shared_ptr<IngestChunk> ch;
        {
            std::lock_guard<std::mutex> lk(mu);
            if (!pool.empty()) { ch = pool.back(); pool.pop_back(); }
        }
        if (!ch) ch = std::make_shared<IngestChunk>();
        ch->tptr = nullptr; ch->tlen = 0; ch->bases.clear(); ch->offs.clear();
        ch->all = ch->skipped = 0; ch->status = 0; ch->err.clear(); ch->parsed = ch->last = false;
        return ch;
    }
    void recycle(std::shared_ptr<IngestChunk> &ch) {
        std::lock_guard<std::mutex> lk(mu);
        if (pool.size() < kMaxQueued + 8) pool.push_back(std::move(ch));
        ch.reset();
    }
    // hands a finished chunk to the workers; false = the reader is being closed
    bool publish(const std::shared_ptr<IngestChunk> &ch) {
        std::unique_lock<std::mutex> lk(mu);
        cv_space.wait(lk, [&] { return stop || ordered.size() < kMaxQueued; });
        if (stop) return false;
        ordered.push_back(ch); todo.push_back(ch);
        cv_work.notify_one();
        return true;
    }

    const uint8_t *map = nullptr; size_t map_len = 0;           // the input file (unmapped when the reader closes)

    void producer() {
        const bool gz = fmt == F_FASTA_GZ || fmt == F_FASTQ_GZ;
        // The file is mapped.  Plain text: the chunks are views of the mapping (no read, no copy; the producer only looks
        // for record boundaries).  .gz: decoded by FastInflate (fast_inflate.h) into pooled buffers.  MFKC_INFLATE=zlib, a
        // file that cannot be mapped, or a .gz name without the gzip magic (gzopen reads those "transparently") go
        // through zlib's gzread / fread as before.
        std::unique_ptr<mfkc::FastInflate> fi;
        const char *inflate_env = getenv("MFKC_INFLATE");
        const bool old_path = inflate_env && !strcmp(inflate_env, "zlib");
        if (!old_path) {
            const int fd = open(path.c_str(), O_RDONLY);
            struct stat sb;
            if (fd >= 0 && fstat(fd, &sb) == 0 && sb.st_size > 0) {
                void *m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                if (m != MAP_FAILED) {
                    if (!gz || mfkc::FastInflate::looks_like_gzip((const uint8_t *)m, (size_t)sb.st_size)) {
                        map = (const uint8_t *)m; map_len = (size_t)sb.st_size;
                        madvise(m, map_len, MADV_SEQUENTIAL);
                    } else munmap(m, (size_t)sb.st_size);
                }
            }
            if (fd >= 0) close(fd);
        }
        if (map && !gz) {                                       // plain text, zero copy
            const char *base = (const char *)map;
            size_t pos = 0;
            bool more = true;
            // FASTQ: finding the cuts is the one serial pass over the text (a 4-line record ends where the newline count
            // is a multiple of 4, so the cut behind a block needs the number of newlines in front of it).  The counting is
            // handed to a few scout threads, block by block; this thread adds the counts up and steps back over
            // (count mod 4) newlines from each block end.  A block with a '\r' or an empty line ends the fast path: the
            // exact line machine takes over at the last cut, as it does for the tail of the file.
            int n_scouts = is_fastq() && map_len > 4 * kChunkText ? std::max(1, std::min(4, n_threads / 3)) : 0;
            if (const char *e = getenv("MFKC_READER_SCOUTS")) n_scouts = is_fastq() ? std::max(0, atoi(e)) : 0;
            if (n_scouts) {
                const size_t C = kChunkText, nb = (map_len + C - 1) / C;
                std::unique_ptr<std::atomic<int64_t>[]> cnt(new std::atomic<int64_t>[nb]);
                for (size_t b = 0; b < nb; b++) cnt[b].store(-2, std::memory_order_relaxed);       // -2 = not counted yet
                std::atomic<size_t> next_block{0};
                std::atomic<bool> quit{false};
                std::vector<std::thread> scouts;
                for (int t = 0; t < n_scouts; t++)
                    scouts.emplace_back([&] {
                        for (;;) {
                            const size_t b = next_block.fetch_add(1);
                            if (b >= nb || quit.load(std::memory_order_relaxed)) return;
                            cnt[b].store(count_clean_newlines(base, b * C, std::min(map_len, (b + 1) * C)), std::memory_order_release);
                        }
                    });
                uint64_t lines = 0;
                bool open_queue = true;
                for (size_t b = 0; b + 1 < nb && open_queue; b++) {               // (the last block belongs to the tail)
                    int64_t c;
                    while ((c = cnt[b].load(std::memory_order_acquire)) == -2) std::this_thread::yield();
                    if (c < 0) break;                                               // not of the simple shape from here on
                    lines += (uint64_t)c;
                    const size_t E = (b + 1) * C;
                    if (E <= pos) continue;
                    const char *q = (const char *)memrchr(base + pos, '\n', E - pos);
                    for (uint64_t r = lines & 3; q && r; r--) q = (const char *)memrchr(base + pos, '\n', (size_t)(q - (base + pos)));
                    if (!q) continue;                                               // no whole record in [pos, E) yet
                    const size_t cut = (size_t)(q - base) + 1;
                    auto ch = fresh_chunk();
                    ch->tptr = base + pos; ch->tlen = cut - pos; ch->last = false;
                    pos = cut;
                    if (!publish(ch)) open_queue = false;
                }
                quit.store(true);
                for (auto &t : scouts) t.join();
                more = open_queue;
            }
            while (more) {
                size_t want = kChunkText, cut = 0, scanned = 0;
                bool eof = false;
                for (;;) {                                      // widen until the block holds at least one whole record
                    const size_t have = std::min(want, map_len - pos);
                    eof = pos + have == map_len;
                    cut = eof ? have : is_fastq() ? last_boundary(base + pos, have, true, false) : fasta_boundary_from(base + pos, scanned, have);
                    if (cut || eof) break;
                    scanned = have;
                    want += kChunkText;
                }
                auto ch = fresh_chunk();
                ch->tptr = base + pos; ch->tlen = cut; ch->last = eof;
                pos += cut;
                more = !eof;
                if (!publish(ch)) break;
            }
            std::lock_guard<std::mutex> lk(mu);
            producer_done = true;
            cv_work.notify_all(); cv_done.notify_all();
            return;
        }
        // big single-member files: several threads decode one stream (parallel_inflate.h); MFKC_INFLATE=serial or a small
        // file: one FastInflate
        std::unique_ptr<mfkc::ParallelInflate> pi;
        std::unique_ptr<mfkc::BgzfInflate> bz;
        if (map && gz) {
            int it = std::max(2, std::min(8, (int)std::thread::hardware_concurrency() * 3 / 4));
            if (const char *e = getenv("MFKC_INFLATE_THREADS")) it = atoi(e);
            if (!(inflate_env && !strcmp(inflate_env, "serial")) && it >= 2) {
                bz.reset(new mfkc::BgzfInflate());                  // bgzip'ed input: member sizes are in the headers
                if (!bz->open(map, map_len, it)) bz.reset();
                if (!bz) {
                    pi.reset(new mfkc::ParallelInflate());
                    if (!pi->open(map, map_len, it)) pi.reset();
                }
            }
            if (!pi && !bz) { fi.reset(new mfkc::FastInflate()); fi->reset(map, map_len); }
        }
        gzFile f = gz && !fi ? gzopen(path.c_str(), "rb") : nullptr;       // plain files: no zlib layer in between
        FILE *pf = gz ? nullptr : fopen(path.c_str(), "rb");
        std::vector<char> carry;
        bool eof = !f && !pf && !fi && !pi && !bz, io_error = eof;
        if (f) gzbuffer(f, 1 << 20);
        while (!eof) {
            auto ch = fresh_chunk();
            if (ch->text.size() < carry.size() + kChunkText) ch->text.resize(carry.size() + kChunkText);
            if (!carry.empty()) memcpy(ch->text.data(), carry.data(), carry.size());
            size_t have = carry.size();
            carry.clear();
            size_t cut = 0, scanned = 0;
            for (;;) {                                          // read until the block holds at least one whole record
                if (ch->text.size() < have + kChunkText) ch->text.resize(have + kChunkText);
                int r;
                if (bz) r = (int)bz->read(ch->text.data() + have, kChunkText);
                else if (pi) r = (int)pi->read(ch->text.data() + have, kChunkText);
                else if (fi) r = (int)fi->read(ch->text.data() + have, kChunkText);
                else if (gz) r = gzread(f, ch->text.data() + have, (unsigned)kChunkText);
                else { r = (int)fread(ch->text.data() + have, 1, kChunkText, pf); if (r == 0 && ferror(pf)) r = -1; }
                if (r < 0) { io_error = true; eof = true; }
                else if (r == 0) eof = true;
                else have += (size_t)r;
                cut = eof ? have : is_fastq() ? last_boundary(ch->text.data(), have, true, false) : fasta_boundary_from(ch->text.data(), scanned, have);
                if (cut || eof) break;
                scanned = have;
            }
            if (!eof) carry.assign(ch->text.data() + cut, ch->text.data() + have);
            ch->tptr = ch->text.data(); ch->tlen = cut;
            ch->last = eof;
            if (io_error) { ch->status = MFKC_E_IO; ch->err = bz && bz->failed() ? "read error (corrupt gzip stream: " + bz->error() + ")" : pi && pi->failed() ? "read error (corrupt gzip stream: " + pi->error() + ")" : fi && fi->failed() ? "read error (corrupt gzip stream: " + fi->error() + ")" : "read error (corrupt gzip stream?)"; }
            if (!publish(ch)) break;
        }
        if (f) gzclose(f);
        if (pf) fclose(pf);
        std::lock_guard<std::mutex> lk(mu);
        producer_done = true;
        cv_work.notify_all(); cv_done.notify_all();
    }

    void worker() {
        for (;;) {
            std::shared_ptr<IngestChunk> ch;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || !todo.empty() || producer_done; });
                if (stop || (todo.empty() && producer_done)) return;
                ch = todo.front(); todo.pop_front();
            }
            RecordParser<MemLineReader> p; p.fmt = fmt; p.phred_lo = phred_lo;
            MemLineReader mlr(ch->tptr, ch->tlen);
            ch->bases.reserve(ch->tlen / 2 + 64);
            ch->offs.push_back(0);
            std::string rd;
            int s;
            while ((s = p.next_read(mlr, rd, &ch->bases)) > 0) ch->offs.push_back(ch->bases.size());
            if (s < 0 && ch->status == 0) { ch->status = s; ch->err = p.err; }
            ch->all = p.all_reads; ch->skipped = p.skipped;
            std::lock_guard<std::mutex> lk(mu);
            ch->parsed = true;
            cv_done.notify_all();
        }
    }

    void start_threads() {
        threads.emplace_back([this] { producer(); });
        for (int i = 0; i < n_threads; i++) threads.emplace_back([this] { worker(); });
    }
    void stop_threads() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_work.notify_all(); cv_space.notify_all(); cv_done.notify_all();
        for (auto &t : threads) t.join();
        threads.clear();
    }
    ~mfkc_reader() {
        if (!threads.empty()) stop_threads();
        if (map) munmap((void *)map, map_len);
    }

    // next parsed chunk in file order into cur_chunk; 1 / 0 (end)
    int next_chunk() {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return (!ordered.empty() && ordered.front()->parsed) || (ordered.empty() && producer_done); });
        if (ordered.empty()) return 0;
        cur_chunk = ordered.front(); ordered.pop_front();
        cur_read = 0;
        cv_space.notify_one();
        return 1;
    }
};

extern "C" int mfkc_reader_open(const char *path, mfkc_reader **out, char *errbuf, size_t err_cap) {
    auto set_err = [&](const std::string &m) { if (errbuf && err_cap) snprintf(errbuf, err_cap, "%s", m.c_str()); };
    if (!path || !out) { set_err("null argument"); return MFKC_E_BADARG; }
    mfkc_reader *r = new mfkc_reader();
    r->path = path;
    r->fmt = detect_format(r->path);
    if (r->fmt == F_UNKNOWN) { set_err("Can't detect file format for file '" + base_name(path) + "'"); delete r; return MFKC_E_FORMAT; }
    if (r->fmt == F_OTHER) { set_err("BINQ / bzip2 inputs are out of scope for libmfkc: " + base_name(path)); delete r; return MFKC_E_FORMAT; }
    r->name = library_name(r->path, r->fmt);
    if (r->is_fastq()) {
        const int s = r->sniff_quality();
        if (s != MFKC_OK) { set_err(r->err); delete r; return s; }
    }
    unsigned nt = std::thread::hardware_concurrency();
    if (const char *e = getenv("MFKC_READER_THREADS")) nt = (unsigned)atoi(e);
    if (nt > 32) nt = 32;
    r->n_threads = nt < 1 ? 1 : (int)nt;
    if (const char *e = getenv("MFKC_READER_CHUNK")) r->kChunkText = std::max<size_t>(16, (size_t)atoll(e));
    if (r->n_threads == 1) {
        if (!r->lr.open(r->path)) { set_err(std::string("can't open ") + path); delete r; return MFKC_E_IO; }
        r->sp.fmt = r->fmt; r->sp.phred_lo = r->phred_lo;
    } else {
        gzFile probe = gzopen(path, "rb");
        if (!probe) { set_err(std::string("can't open ") + path); delete r; return MFKC_E_IO; }
        gzclose(probe);
        r->start_threads();
    }
    *out = r;
    return MFKC_OK;
}

// NamedSource.name() of readDnaLazy(file) without opening a reader (no threads, no sniff, no I/O beyond nothing at all)
extern "C" int mfkc_library_name(const char *path, char *out, size_t cap) {
    if (!path || !out || !cap) return MFKC_E_BADARG;
    const Format f = detect_format(path);
    if (f == F_UNKNOWN || f == F_OTHER) { out[0] = 0; return MFKC_E_FORMAT; }
    snprintf(out, cap, "%s", library_name(path, f).c_str());
    return MFKC_OK;
}

extern "C" int mfkc_reader_pending_bases(const mfkc_reader *r, uint64_t *n_bases) {
    if (!r || !n_bases) return MFKC_E_BADARG;
    *n_bases = 0;
    if (r->n_threads > 1) {
        if (r->cur_chunk && r->cur_read + 1 < r->cur_chunk->offs.size())
            *n_bases = r->cur_chunk->offs[r->cur_read + 1] - r->cur_chunk->offs[r->cur_read];
    } else if (r->have_cur) *n_bases = r->cur.size();
    return MFKC_OK;
}

extern "C" int mfkc_reader_next(mfkc_reader *r, uint8_t *bases, size_t cap_bases, uint64_t *offsets, uint32_t cap_reads,
                                uint32_t *n_reads) {
    if (!r || !bases || !offsets || !n_reads) return MFKC_E_BADARG;
    uint32_t n = 0;
    size_t used = 0;
    offsets[0] = 0;
    if (r->n_threads > 1) {
        while (n < cap_reads && !r->done) {
            if (!r->cur_chunk) {
                if (r->next_chunk() == 0) { r->done = true; break; }
                r->all_reads += r->cur_chunk->all; r->skipped += r->cur_chunk->skipped;
            }
            IngestChunk &c = *r->cur_chunk;
            const size_t n_in = c.offs.size() - 1;
            bool full = false;
            // whole runs of reads at once: the offsets of a chunk are already a prefix sum
            while (r->cur_read < n_in && n < cap_reads) {
                size_t take = std::min<size_t>(n_in - r->cur_read, cap_reads - n);
                const uint64_t b0 = c.offs[r->cur_read];
                while (take && c.offs[r->cur_read + take] - b0 > cap_bases - used) take = take > 64 ? take / 2 : take - 1;
                if (!take) {
                    if (n == 0) { r->err = "read longer than the batch buffer"; return MFKC_E_BADARG; }
                    full = true; break;
                }
                const uint64_t nb = c.offs[r->cur_read + take] - b0;
                memcpy(bases + used, c.bases.data() + b0, nb);
                for (size_t i = 1; i <= take; i++) offsets[n + i] = used + (c.offs[r->cur_read + i] - b0);
                used += nb; n += (uint32_t)take; r->cur_read += take;
            }
            if (full || n == cap_reads) break;
            if (c.status < 0) {                              // the error sits behind the reads of this chunk
                if (n) break;                                // hand the reads out first; the error comes with the next call
                r->err = c.err; r->done = true; *n_reads = 0;
                return c.status;
            }
            if (c.last) { r->done = true; }
            r->recycle(r->cur_chunk);
        }
        *n_reads = n;
        return MFKC_OK;
    }
    while (n < cap_reads) {
        if (!r->have_cur) {
            const int s = r->advance();
            if (s < 0) { *n_reads = n; return s; }
            if (s == 0) break;
            r->have_cur = true;
        }
        if (used + r->cur.size() > cap_bases) {
            if (n == 0) { r->err = "read longer than the batch buffer"; return MFKC_E_BADARG; }
            break;
        }
        memcpy(bases + used, r->cur.data(), r->cur.size());
        used += r->cur.size();
        offsets[++n] = used;
        r->have_cur = false;
    }
    *n_reads = n;
    return MFKC_OK;
}

extern "C" int mfkc_reader_counters(const mfkc_reader *r, uint64_t counters[2]) {
    if (!r || !counters) return MFKC_E_BADARG;
    counters[0] = r->all_reads; counters[1] = r->skipped;
    return MFKC_OK;
}
extern "C" const char *mfkc_reader_error(const mfkc_reader *r) { return r ? r->err.c_str() : ""; }
extern "C" const char *mfkc_reader_name(const mfkc_reader *r) { return r ? r->name.c_str() : ""; }
extern "C" void mfkc_reader_close(mfkc_reader *r) { delete r; }

// QuickQuantitativeStatistics.printToFile (:65-72): println(header); println(toString())
extern "C" int mfkc_write_stat_file(const char *path, const uint64_t hist[MFKC_HIST_BINS]) {
    if (!path || !hist) return MFKC_E_BADARG;
    FILE *f = fopen(path, "w");
    if (!f) return MFKC_E_IO;
    fputs("# k-mer frequency\tnumber of such k-mers\n", f);           // src/io/IOUtils.java:69
    for (int c = 0; c < MFKC_HIST_BINS; c++)
        if (hist[c]) fprintf(f, "%d\t%llu\n", c, (unsigned long long)hist[c]);
    fputs("\n", f);
    return fclose(f) == 0 ? MFKC_OK : MFKC_E_IO;
}

// ------------------------------------------------------------------------------------------
// synthetic reads: community tables + host generator
// ------------------------------------------------------------------------------------------
extern "C" void mfkc_synth_defaults(mfkc_synth_cfg *c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
    c->struct_size = sizeof *c;
    c->n_genomes = 64;
    c->seed = 0x4D464B43ULL;                 // "MFKC"
    c->total_genome_bp = 150000000ULL;
    c->read_len = 150;
    c->sample = 0;
    c->err_ppm_first = 1000; c->err_ppm_last = 10000;
    c->n_read_ppm = 1000; c->poly_tail_ppm = 100;
}

static double u01(uint64_t x) { return ((x >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

// Genome lengths log-uniform in [0.5, 8] Mbp rescaled to total_genome_bp; abundances
// log-normal(sigma = 1) per (seed, sample); read-sampling weight = abundance x length.
extern "C" int mfkc_synth_tables_build(const mfkc_synth_cfg *c, mfkc_synth_tables *t) {
    if (!c || !t || c->struct_size != sizeof(mfkc_synth_cfg)) return MFKC_E_BADARG;
    if (c->n_genomes == 0 || c->n_genomes > MFKC_SYNTH_MAX_GENOMES || c->read_len == 0 || c->err_ppm_last < c->err_ppm_first)
        return MFKC_E_BADARG;
    const uint32_t G = c->n_genomes;
    memset(t, 0, sizeof *t);
    t->seed = c->seed; t->n_genomes = G; t->read_len = c->read_len; t->sample = c->sample;
    t->err_ppm_first = c->err_ppm_first; t->err_ppm_last = c->err_ppm_last;
    t->n_read_ppm = c->n_read_ppm; t->poly_tail_ppm = c->poly_tail_ppm;
    std::vector<double> len(G), w(G);
    double sum = 0;
    for (uint32_t g = 0; g < G; g++) {
        const double u = u01(mfkc_splitmix64(c->seed ^ 0x4C454E00ULL ^ ((uint64_t)g << 32)));
        len[g] = std::exp(std::log(0.5e6) + u * std::log(16.0));
        sum += len[g];
    }
    uint64_t off = 0;
    for (uint32_t g = 0; g < G; g++) {
        uint64_t L = (uint64_t)(len[g] * ((double)c->total_genome_bp / sum));
        if (L < 2ull * c->read_len) L = 2ull * c->read_len;
        t->genome_off[g] = off;
        off += L;
        const uint64_t h1 = mfkc_splitmix64(c->seed ^ 0x41424E44ULL ^ ((uint64_t)c->sample << 40) ^ ((uint64_t)g << 8));
        const uint64_t h2 = mfkc_splitmix64(h1);
        const double z = std::sqrt(-2.0 * std::log(u01(h1))) * std::cos(6.283185307179586 * u01(h2));   // Box-Muller
        w[g] = std::exp(z) * (double)L;
    }
    t->genome_off[G] = off;
    double W = 0, run = 0;
    for (uint32_t g = 0; g < G; g++) W += w[g];
    for (uint32_t g = 0; g < G; g++) {
        run += w[g];
        const double frac = run / W;
        t->cum_weight[g] = frac >= 1.0 ? ~0ULL : (uint64_t)(frac * 18446744073709551616.0);
    }
    t->cum_weight[G - 1] = ~0ULL;
    return MFKC_OK;
}

extern "C" int mfkc_synth_reads_host(const mfkc_synth_cfg *cfg, uint64_t first_read, uint64_t n_reads, uint8_t *out) {
    if (!cfg || (!out && n_reads)) return MFKC_E_BADARG;
    mfkc_synth_tables *t = new mfkc_synth_tables();
    const int r = mfkc_synth_tables_build(cfg, t);
    if (r != MFKC_OK) { delete t; return r; }
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 64) nt = 64;
    if (n_reads < 4096) nt = 1;
    std::vector<std::thread> th;
    const uint32_t L = t->read_len;
    for (unsigned w = 0; w < nt; w++) {
        const uint64_t lo = n_reads * w / nt, hi = n_reads * (w + 1) / nt;
        th.emplace_back([=]() {
            for (uint64_t i = lo; i < hi; i++) {
                mfkc_synth_read rd;
                mfkc_synth_read_header(*t, first_read + i, rd);
                uint8_t *o = out + i * L;
                for (uint32_t j = 0; j < L; j++) o[j] = mfkc_synth_read_base(*t, rd, j);
            }
        });
    }
    for (auto &x : th) x.join();
    delete t;
    return MFKC_OK;
}

// ------------------------------------------------------------------------------------------
// Deterministic merge of per-shard record streams (multi-GPU runs, SURVEY.md 8e: "a deterministic sorted
// merge produces the .kmers.bin output").  Every part is sorted by ascending big-endian key and the parts hold
// disjoint key sets (one owner shard per k-mer), so the merge is an interleave; equal keys (never produced by the
// shards) would keep part order.  The key range is cut at sampled splitters and every host thread merges one slice
// of all parts into its place of `out`, so the result does not depend on the thread count.
// ------------------------------------------------------------------------------------------
namespace {
struct MergePart { const uint8_t *p; uint64_t n; };

inline int key_cmp(const uint8_t *a, const uint8_t *b, uint32_t key_bytes) { return memcmp(a, b, key_bytes); }

// first record of part with key >= key
uint64_t lower_bound_rec(const MergePart &part, const uint8_t *key, uint32_t rs, uint32_t kb) {
    uint64_t lo = 0, hi = part.n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (key_cmp(part.p + mid * rs, key, kb) < 0) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// big-endian key of a record as a native integer (8-byte keys: k <= 31; 16-byte keys: k > 31)
template <class K> inline K load_key(const uint8_t *p);
template <> inline uint64_t load_key<uint64_t>(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return __builtin_bswap64(v); }
typedef unsigned __int128 u128_t;
template <> inline u128_t load_key<u128_t>(const uint8_t *p) {
    uint64_t hi, lo; memcpy(&hi, p, 8); memcpy(&lo, p + 8, 8);
    return ((u128_t)__builtin_bswap64(hi) << 64) | __builtin_bswap64(lo);
}

// One slice of the key range: records [b[i], e[i]) of every part into `out`.  The shards split the keys by a hash, so
// the parts interleave record by record and the cost is what is spent per RECORD: the heads' keys live in a small
// array as native integers, the smallest is found by a branch-free scan (compare + conditional moves; ties go to the
// lowest part, so equal keys -- never produced by the shards -- would keep part order), the record moves with two
// fixed-size copies.  An exhausted part leaves the array (order kept); the last part standing is one memcpy.
// One slice of the key range: records [b[i], e[i]) of every part into `out`.  The shards split the keys by a hash, so
// the parts interleave record by record and the cost is what is spent per RECORD: the heads' keys live in a small
// array as native integers, the smallest is found by a branch-free scan (compare + conditional moves; ties go to the
// lowest part, so equal keys -- never produced by the shards -- would keep part order), the record moves with two
// fixed-size copies.  An exhausted part leaves the array (order kept); the last part standing is one memcpy.
// (Tried and dropped, both slower here: a balanced compare tree over 8 / 16 padded slots, and register-resident heads
// with the next key loaded one pop ahead.)
template <class K, uint32_t RS>
void merge_slice_t(const std::vector<MergePart> &parts, const std::vector<uint64_t> &b, const std::vector<uint64_t> &e, uint8_t *out) {
    constexpr uint32_t KB = RS - 2;
    const size_t P = parts.size();
    std::vector<K> hk(P);
    std::vector<const uint8_t *> hp(P), he(P);
    size_t n = 0;
    for (size_t i = 0; i < P; i++)
        if (b[i] < e[i]) { hp[n] = parts[i].p + b[i] * RS; he[n] = parts[i].p + e[i] * RS; hk[n] = load_key<K>(hp[n]); n++; }
    K *k = hk.data();
    while (n > 1) {
        size_t best = 0; K bk = k[0];
        for (size_t i = 1; i < n; i++) { const bool lt = k[i] < bk; bk = lt ? k[i] : bk; best = lt ? i : best; }
        const uint8_t *src = hp[best];
        memcpy(out, src, KB); memcpy(out + KB, src + KB, 2);
        out += RS; src += RS;
        hp[best] = src;
        if (__builtin_expect(src != he[best], 1)) k[best] = load_key<K>(src);
        else {
            for (size_t i = best + 1; i < n; i++) { k[i - 1] = k[i]; hp[i - 1] = hp[i]; he[i - 1] = he[i]; }
            n--;
        }
    }
    if (n == 1) memcpy(out, hp[0], (size_t)(he[0] - hp[0]));
}

}  // namespace

extern "C" int mfkc_merge_records(const uint8_t *const *parts_in, const uint64_t *n_records, uint32_t n_parts, uint32_t record_size,
                                  uint8_t *out, int threads) {
    if ((!parts_in || !n_records) && n_parts) return MFKC_E_BADARG;
    if (record_size != 10 && record_size != 18) return MFKC_E_BADARG;
    const uint32_t kb = record_size - 2;
    std::vector<MergePart> parts;
    uint64_t total = 0, biggest = 0; size_t big_i = 0;
    for (uint32_t i = 0; i < n_parts; i++) {
        if (n_records[i] && !parts_in[i]) return MFKC_E_BADARG;
        parts.push_back({parts_in[i], n_records[i]});
        total += n_records[i];
        if (n_records[i] > biggest) { biggest = n_records[i]; big_i = i; }
    }
    if (!total) return MFKC_OK;
    if (!out) return MFKC_E_BADARG;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    uint32_t T = (uint32_t)std::min<uint64_t>((uint64_t)std::min(threads, 64), std::max<uint64_t>(1, total >> 16));
    // slice j = keys in [splitter[j-1], splitter[j]); splitters = quantiles of the biggest part
    std::vector<std::vector<uint64_t>> cut(T + 1, std::vector<uint64_t>(parts.size()));
    for (size_t i = 0; i < parts.size(); i++) { cut[0][i] = 0; cut[T][i] = parts[i].n; }
    for (uint32_t j = 1; j < T; j++) {
        const uint8_t *sk = parts[big_i].p + (biggest * j / T) * record_size;
        for (size_t i = 0; i < parts.size(); i++) cut[j][i] = std::max(cut[j - 1][i], lower_bound_rec(parts[i], sk, record_size, kb));
    }
    std::vector<uint64_t> out_off(T + 1, 0);
    for (uint32_t j = 1; j <= T; j++) { uint64_t s = 0; for (size_t i = 0; i < parts.size(); i++) s += cut[j][i]; out_off[j] = s; }
    std::vector<std::thread> th;
    for (uint32_t j = 0; j < T; j++) {
        auto work = [&, j] {
            uint8_t *o = out + out_off[j] * record_size;
            if (record_size == 10) merge_slice_t<uint64_t, 10>(parts, cut[j], cut[j + 1], o);
            else merge_slice_t<u128_t, 18>(parts, cut[j], cut[j + 1], o);
        };
        if (T == 1) work(); else th.emplace_back(work);
    }
    for (auto &t : th) t.join();
    return MFKC_OK;
}
