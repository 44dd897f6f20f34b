// host_io.cpp -- CPU side of the path behind the same C ABI (no GPU needed):
//   * FASTA / FASTQ (+ .gz) readers with the reference's parser rules
//       [itmo]/io/ReadersUtils.java:27-54,57-77,85-102
//       [itmo]/io/readers/FastaReader.java:54-108, FastqReader.java:53-114,
//       [itmo]/io/readers/FastaReaderFromXQSource.java:62-85,
//       [itmo]/io/formats/Illumina.java:7-12, Sanger.java:7-12, [itmo]/dna/DnaQ.java:140-150
//   * .stat.txt writer ([itmo]/statistics/QuickQuantitativeStatistics.java:38-72)
//   * the synthetic read generator's host half (synth.h)
#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mfkc.h"
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

// BufferedReader.readLine over a (possibly gzip) stream: terminators \n, \r\n, \r.
class LineReader {
public:
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
            size_t i = 0;
            while (i < avail && p[i] != '\n' && p[i] != '\r') i++;
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

struct mfkc_reader {
    std::string path, name, err;
    Format fmt = F_UNKNOWN;
    LineReader lr;
    int phred_lo = 64;                     // Illumina (64) unless the sniff says Sanger (33)
    uint64_t all_reads = 0, skipped = 0;
    bool done = false;
    std::string cur;                       // next kept read, not yet handed out
    bool have_cur = false;
    // FASTA record assembly
    std::string sb;
    bool fasta_eof = false;

    // ---- FASTA: FastaReader.java:82-104 (readNextDataLine) + :54-66 (N-drop)
    // returns 1 = record in `out`, 0 = end, <0 error
    int fasta_next(std::string &out) {
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
    int fastq_data_line(LineReader &r, const char *&l, size_t &n) {
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
    int fastq_record(LineReader &r, int lo, std::string *out, bool &kept) {
        const char *l; size_t n;
        int s = fastq_data_line(r, l, n);
        if (s <= 0) return s;
        data_.assign(l, n);
        s = fastq_data_line(r, l, n);
        if (s == 0) { err = "Unexpected end of file. File is corrupted/Format mismatch."; return MFKC_E_FORMAT; }
        if (s < 0) return s;
        if (n != data_.size()) {
            err = "Bad DnaQ record: length of chars and quality is not the same. File is corrupted/Format mismatch.";
            return MFKC_E_FORMAT;
        }
        kept = true;
        for (size_t i = 0; i < n; i++) {
            const unsigned char ch = (unsigned char)data_[i];
            if (ch == 'N' || ch == 'n' || ch == '.') { kept = false; continue; }    // appendUnknown: phred 0
            if (!kCode.ok[ch]) { err = std::string("Incorrect nucleotide char: \"") + (char)ch + "\""; return MFKC_E_FORMAT; }
            const int q = (unsigned char)l[i];
            if (q < lo || q > 126) {                                                // Illumina.java:8 / Sanger.java:8
                err = std::string("Invalid quality code char: \"") + (char)q + "\"";
                return -100;
            }
            if (((q - lo) & 63) == 0) kept = false;       // 6-bit phred (DnaQ.java:140-150) == 0 -> read dropped
        }
        if (out && kept) out->swap(data_);
        return 1;
    }

    // ReadersUtils.determineQualityFormat :63-77
    int sniff_quality() {
        LineReader r;
        if (!r.open(path)) { err = "can't open " + path; return MFKC_E_IO; }
        for (int i = 0; i < 1000; i++) {
            bool kept;
            const int s = fastq_record(r, 64, nullptr, kept);
            if (s == 0) break;
            if (s == -100) { phred_lo = 33; err.clear(); return MFKC_OK; }
            if (s < 0) return s;
        }
        phred_lo = 64;
        return MFKC_OK;
    }

    // next kept read into cur; 1 / 0 / <0
    int advance() {
        if (done) return 0;
        int s;
        if (fmt == F_FASTA || fmt == F_FASTA_GZ) s = fasta_next(cur);
        else {
            for (;;) {
                bool kept = false;
                s = fastq_record(lr, phred_lo, &cur, kept);
                if (s == -100) s = MFKC_E_FORMAT;
                if (s <= 0) break;
                all_reads++;
                if (kept) break;
                skipped++;
            }
            if (s == 0 && lr.failed()) { err = "read error (corrupt gzip stream?)"; s = MFKC_E_IO; }
        }
        if (s <= 0) done = true;
        return s;
    }

private:
    std::string data_;
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
    if (r->fmt == F_FASTQ || r->fmt == F_FASTQ_GZ) {
        const int s = r->sniff_quality();
        if (s != MFKC_OK) { set_err(r->err); delete r; return s; }
    }
    if (!r->lr.open(r->path)) { set_err(std::string("can't open ") + path); delete r; return MFKC_E_IO; }
    *out = r;
    return MFKC_OK;
}

extern "C" int mfkc_reader_next(mfkc_reader *r, uint8_t *bases, size_t cap_bases, uint64_t *offsets, uint32_t cap_reads,
                                uint32_t *n_reads) {
    if (!r || !bases || !offsets || !n_reads) return MFKC_E_BADARG;
    uint32_t n = 0;
    size_t used = 0;
    offsets[0] = 0;
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
