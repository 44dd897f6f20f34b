// cli.cpp -- mfkc_cli: the host side of the path as a native tool with MetaFast's own command
// lines, for boxes without a JVM (tests, benchmarks) and as the executable specification of what
// the Java Tool subclasses do on top of the C ABI (java/ holds those classes).
//
//   mfkc_cli -t kmer-counter-many  -k K [-b B] -i reads... [-w workDir] [--output-dir D] [--stats-dir D]
//   mfkc_cli -t kmer-counter       -k K [-b B] -i reads...            (one sample: all files into one table)
//   mfkc_cli -t features-calculator -k K -cm components.bin [-ka kmers.bin...] [-i reads...]
//                                   [--selected kmers.bin...] [--threshold T] [-w workDir]
//   mfkc_cli -t seq-builder | seq-builder-many | component-cutter | dist-matrix-calculator | heatmap-maker | matrix-builder ...
//   mfkc_cli -t kmers-filter | unique-kmers-multi | kmers-samples-counter ...        
//   extras (optional, old command lines keep working): --gpu N, --gpu-variant hash|sort|direct,
//   --long-kmers (accept 32 <= k <= 63: 128-bit keys, <name>.kmers128.bin with 18-byte records)
//
// Mirrors src/tools/KmersCounterForManyFilesMain.java:26-120, src/tools/KmersCounterMain.java:28-137,
// src/tools/FeaturesCalculatorMain.java:30-236 and src/structures/ConnectedComponent.java:95-122:
// same option names, defaults, output locations, log lines and exit codes.
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <memory>
#include <vector>

#include "../../include/mfkc.h"

namespace {

[[noreturn]] void die(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    fputs("ERROR: ", stderr); vfprintf(stderr, fmt, ap); fputc('\n', stderr);
    va_end(ap);
    // Tool.java:450-462: logged, System.exit(1).  _exit: reader / worker threads of other open files and GPU contexts may
    // still be running; static destructors and CUDA teardown under their feet could crash and change the exit code
    fflush(stdout); fflush(stderr);
    _exit(1);
}
void info(const char *fmt, ...) { va_list ap; va_start(ap, fmt); fputs("INFO: ", stderr); vfprintf(stderr, fmt, ap); fputc('\n', stderr); va_end(ap); }
void warn(const char *fmt, ...) { va_list ap; va_start(ap, fmt); fputs("WARN: ", stderr); vfprintf(stderr, fmt, ap); fputc('\n', stderr); va_end(ap); }

// NumUtils.groupDigits ([itmo]/utils/NumUtils.java:163-174): 1'234'567
std::string group_digits(unsigned long long v) {
    std::string vs = std::to_string(v), ans;
    while (vs.size() > 3) { ans = "'" + vs.substr(vs.size() - 3) + ans; vs.resize(vs.size() - 3); }
    return vs + ans;
}

void mkdirs(const std::string &p) {
    std::string cur;
    for (size_t i = 0; i <= p.size(); i++) {
        if (i == p.size() || p[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0777); }
        if (i < p.size()) cur += p[i];
    }
}
std::string base_name(const std::string &p) { const size_t i = p.find_last_of('/'); return i == std::string::npos ? p : p.substr(i + 1); }
bool ends_with(const std::string &s, const std::string &suf) { return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0; }
bool ends_with_ci(const std::string &s, const std::string &suf) {
    return s.size() >= suf.size() && strcasecmp(s.c_str() + s.size() - suf.size(), suf.c_str()) == 0;
}
long long file_size(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0 ? (long long)st.st_size : 0; }

// java.lang.Double.toString (JDK >= 19: shortest digits that round-trip)
std::string java_double(double d) {
    if (std::isnan(d)) return "NaN";
    if (std::isinf(d)) return d > 0 ? "Infinity" : "-Infinity";
    if (d == 0) return std::signbit(d) ? "-0.0" : "0.0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(d), std::chars_format::scientific);   // d.ddddde[+-]xx, shortest
    std::string s(buf, r.ptr);
    const size_t e = s.find('e');
    std::string digits = s.substr(0, e);
    const int exp10 = atoi(s.c_str() + e + 1);
    digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const std::string sign = d < 0 ? "-" : "";
    const double a = std::fabs(d);
    if (a >= 1e-3 && a < 1e7) {
        const int point = exp10 + 1;                     // digits before the decimal point
        std::string out;
        if (point <= 0) out = "0." + std::string(-point, '0') + digits;
        else if ((size_t)point >= digits.size()) out = digits + std::string(point - digits.size(), '0') + ".0";
        else out = digits.substr(0, point) + "." + digits.substr(point);
        return sign + out;
    }
    std::string frac = digits.size() > 1 ? digits.substr(1) : "0";
    return sign + digits.substr(0, 1) + "." + frac + "E" + std::to_string(exp10);
}

std::atomic<int> g_ctx_gen_counter{0};
thread_local int g_ctx_gen = 0;                           // set by make_ctx in the calling thread: function-local caches of pinned buffers follow it
bool g_sub_tool = false;                                  // inside matrix-builder: sub-tools do not print their output values
#define OUT_VALUE(...) do { if (!g_sub_tool) printf(__VA_ARGS__); } while (0)

struct Args {
    std::string tool;
    std::map<std::string, std::vector<std::string>> opt;    // canonical long name -> values
    bool has(const std::string &k) const { return opt.count(k) != 0; }
    std::string one(const std::string &k, const std::string &def = "") const {
        auto it = opt.find(k);
        return it == opt.end() || it->second.empty() ? def : it->second[0];
    }
    std::vector<std::string> many(const std::string &k) const { auto it = opt.find(k); return it == opt.end() ? std::vector<std::string>() : it->second; }
};

// commons-cli PosixParser behaviour that matters here: a multi-valued option takes every following
// token up to the next option ([itmo]/utils/tool/parameters/MultiValuedParameter.java:13-18).
Args parse_args(int argc, char **argv) {
    static const std::map<std::string, std::string> alias = {
        {"-t", "tool"}, {"--tool", "tool"}, {"-k", "k"}, {"--k", "k"}, {"-i", "reads"}, {"--reads", "reads"},
        {"-b", "maximal-bad-frequence"}, {"--maximal-bad-frequence", "maximal-bad-frequence"},
        {"--output-dir", "output-dir"}, {"--stats-dir", "stats-dir"}, {"-w", "work-dir"}, {"--work-dir", "work-dir"},
        {"-cm", "components-file"}, {"--components-file", "components-file"}, {"-ka", "kmers"}, {"--kmers", "kmers"},
        {"--selected", "selected"}, {"--threshold", "threshold"}, {"-p", "available-processors"},
        {"--available-processors", "available-processors"}, {"--gpu", "gpu"}, {"--gpu-variant", "gpu-variant"},
        {"--gpus", "gpus"}, {"--gpu-mode", "gpu-mode"},
        {"--force", "force"}, {"-v", "verbose"}, {"--verbose", "verbose"}, {"--long-kmers", "long-kmers"},
        {"--k-mers", "reads"}, {"--filter-kmers", "filter-kmers"}, {"--max-thresh", "max-thresh"},
        {"--min-samples", "min-samples"}, {"--max-samples", "max-samples"}, {"--min-seq-len", "min-seq-len"}, {"-l", "l"},
        {"--maximal-bad-frequency", "maximal-bad-frequence"}, {"-bp", "bottom-cut-percent"}, {"--bottom-cut-percent", "bottom-cut-percent"},
        {"--sequence-len", "sequence-len"}, {"-o", "output-dir"},
        {"--sequences", "reads"}, {"-b1", "min-component-size"}, {"--min-component-size", "min-component-size"},
        {"-b2", "max-component-size"}, {"--max-component-size", "max-component-size"}, {"--features", "features"},
        {"-wn", "without-names"}, {"--without-names", "without-names"}, {"--matrix-file", "matrix-file"},
        {"--output-format", "output-format"}, {"--newMatrix-file", "newMatrix-file"}, {"-wr", "without-renumbering"},
        {"--without-renumbering", "without-renumbering"}, {"--heatmap-file", "heatmap-file"},
        {"--use-reads-for-calculating-features", "use-reads-for-calculating-features"}};
    Args a;
    std::string cur;
    for (int i = 1; i < argc; i++) {
        const std::string t = argv[i];
        const bool is_opt = t.size() > 1 && t[0] == '-' && !(t[1] >= '0' && t[1] <= '9');
        if (is_opt) {
            auto it = alias.find(t);
            if (it == alias.end()) die("Unrecognized option: %s", t.c_str());
            cur = it->second;
            a.opt[cur];
        } else {
            if (cur.empty()) die("Unexpected argument: %s", t.c_str());
            a.opt[cur].push_back(t);
        }
    }
    a.tool = a.one("tool", "matrix-builder");
    return a;
}

int parse_int(const Args &a, const std::string &key, bool mandatory, int def) {
    if (!a.has(key) || a.many(key).empty()) {
        if (mandatory) die("Missing mandatory parameter --%s", key.c_str());
        return def;
    }
    char *end = nullptr;
    const std::string v = a.one(key);
    const long x = strtol(v.c_str(), &end, 10);
    if (*end) die("Can't parse '%s' as integer for --%s", v.c_str(), key.c_str());
    return (int)x;
}

struct Gpu {
    int device = 0;               // first device
    int variant = MFKC_VARIANT_HASH;
    int n_gpus = 1;               // --gpus G: devices device .. device + G - 1
    std::string mode = "auto";    // --gpu-mode samples (one sample per GPU at a time) | shard (every sample over all GPUs) | auto
    int n_shards = 0, shard_id = 0;
};
Gpu gpu_opts(const Args &a) {
    Gpu g;
    g.device = parse_int(a, "gpu", false, 0);
    const std::string v = a.one("gpu-variant", "hash");
    if (v == "hash") g.variant = MFKC_VARIANT_HASH;
    else if (v == "sort") g.variant = MFKC_VARIANT_SORT;
    else if (v == "direct") g.variant = MFKC_VARIANT_HASH_DIRECT;
    else if (v == "table") g.variant = MFKC_VARIANT_HASH_TABLE;
    else die("--gpu-variant must be hash, table, sort or direct");
    g.n_gpus = parse_int(a, "gpus", false, 1);
    if (g.n_gpus < 1 || g.n_gpus > 16) die("--gpus must be 1..16");
    g.mode = a.one("gpu-mode", "auto");
    if (g.mode != "auto" && g.mode != "samples" && g.mode != "shard") die("--gpu-mode must be auto, samples or shard");
    if (g.n_gpus > 1 && g.device + g.n_gpus > mfkc_device_count() && !getenv("MFKC_LOGICAL_GPUS"))
        die("--gpus %d: only %d CUDA device(s) visible", g.n_gpus, mfkc_device_count());
    return g;
}

std::string reader_name(const std::string &path) {
    char name[4096] = "";
    if (mfkc_library_name(path.c_str(), name, sizeof name) != MFKC_OK) {
        // same message and moment as the reference: opening the file fails (ReadersUtils.readDnaLazy)
        mfkc_reader *r = nullptr; char err[512] = "";
        if (mfkc_reader_open(path.c_str(), &r, err, sizeof err) != MFKC_OK) die("%s", err);
        std::string n = mfkc_reader_name(r);
        mfkc_reader_close(r);
        return n;
    }
    return name;
}

#define CK(ctx, call) do { int rc__ = (call); if (rc__ != MFKC_OK) die("%s (libmfkc %d)", mfkc_last_error(ctx), rc__); } while (0)

// ---- IOUtils.loadReads for one file: stream batches of kept reads into `submit`
mfkc_reader *open_reader(const std::string &file) {
    mfkc_reader *r = nullptr; char err[512] = "";
    if (mfkc_reader_open(file.c_str(), &r, err, sizeof err) != MFKC_OK) die("%s", err);
    return r;
}
// `r`: a reader opened earlier (its producer / parser threads have been running since), or nullptr
template <class F>
void for_each_batch(mfkc_ctx *ctx, const std::string &file, F submit, mfkc_reader *r = nullptr) {
    info("Loading file %s...", base_name(file).c_str());
    if (!r) r = open_reader(file);
    const uint32_t cap_reads = 1u << 21;
    static thread_local void *h_bases = nullptr, *h_offs = nullptr; static thread_local int owner_gen = -1;      // pinned buffers die with their context
    static thread_local size_t cap_bases = 0;
    if (owner_gen != g_ctx_gen) { h_bases = h_offs = nullptr; owner_gen = g_ctx_gen; }
    if (!h_bases) {
        cap_bases = 256u << 20;
        if (const char *e = getenv("MFKC_CLI_BATCH_BASES")) cap_bases = std::max<size_t>(64, strtoull(e, nullptr, 10));      // small batches: tests
        CK(ctx, mfkc_pinned_alloc(ctx, cap_bases, &h_bases)); CK(ctx, mfkc_pinned_alloc(ctx, ((size_t)cap_reads + 1) * 8, &h_offs));
    }
    unsigned long long reads = 0;
    for (;;) {
        uint32_t n = 0;
        int rc = mfkc_reader_next(r, (uint8_t *)h_bases, cap_bases, (uint64_t *)h_offs, cap_reads, &n);
        uint64_t pending = 0;
        if (rc == MFKC_E_BADARG && mfkc_reader_pending_bases(r, &pending) == MFKC_OK && pending > cap_bases) {
            // one record longer than the batch buffer (a chromosome-sized FASTA record; the reference takes any length):
            // the reader kept it, take it with a larger buffer
            (void)mfkc_pinned_free(ctx, h_bases);      // (a buffer that another context of this thread allocated stays with it until it is destroyed)
            h_bases = nullptr;
            cap_bases = (size_t)pending + (pending >> 3);
            CK(ctx, mfkc_pinned_alloc(ctx, cap_bases, &h_bases));
            rc = mfkc_reader_next(r, (uint8_t *)h_bases, cap_bases, (uint64_t *)h_offs, cap_reads, &n);
        }
        if (rc != MFKC_OK) die("%s: %s", file.c_str(), mfkc_reader_error(r));
        if (!n) break;
        submit((const uint8_t *)h_bases, (const uint64_t *)h_offs, n);
        reads += n;
    }
    uint64_t c[2]; mfkc_reader_counters(r, c);
    if (c[1]) info("Skipped %s (%.1f%%) out of %s reads (because of N nucleotide), file %s", group_digits(c[1]).c_str(),
                   c[1] * 100.0 / c[0], group_digits(c[0]).c_str(), base_name(file).c_str());
    info("%s reads added", group_digits(reads).c_str());               // src/io/IOUtils.java:863
    mfkc_reader_close(r);
}

void report_sample(int k, uint64_t size, uint64_t good, const std::string &out_file);

// ---- KmersCounterMain.runImpl (src/tools/KmersCounterMain.java:65-120) for one sample
std::string count_sample(mfkc_ctx *ctx, int k, int b, const std::vector<std::string> &files, const std::string &name,
                         const std::string &out_dir, const std::string &st_dir) {
    CK(ctx, mfkc_reset(ctx));
    // a sliding window of two open readers: file i + 1 inflates and parses in the background (mfkc_reader_open starts its
    // threads) while file i is being submitted, whatever the number of files of the sample (a reader holds dozens of
    // threads and tens of MB of queued chunks)
    mfkc_reader *next = files.empty() ? nullptr : open_reader(files[0]);
    for (size_t i = 0; i < files.size(); i++) {
        mfkc_reader *cur = next;
        next = i + 1 < files.size() ? open_reader(files[i + 1]) : nullptr;
        for_each_batch(ctx, files[i], [&](const uint8_t *bases, const uint64_t *offs, uint32_t n) { CK(ctx, mfkc_submit_reads(ctx, bases, offs, n)); }, cur);
    }
    CK(ctx, mfkc_flush(ctx));
    mkdirs(out_dir); mkdirs(st_dir);
    // k > 31 (--long-kmers, 18-byte records) gets its own extension so that no reference tool misreads the file
    const std::string out_file = out_dir + "/" + name + (k > 31 ? ".kmers128.bin" : ".kmers.bin"), st_file = st_dir + "/" + name + ".stat.txt";
    uint64_t good = 0;
    CK(ctx, mfkc_emit_begin(ctx, b, &good));
    FILE *f = fopen(out_file.c_str(), "wb");
    if (!f) die("Can't write %s", out_file.c_str());
    static thread_local void *h_out = nullptr; const size_t chunk = 16777200;       // KMERS_WORK_RANGE_SIZE, src/io/IOUtils.java:30
    static thread_local int owner_gen = -1;
    if (owner_gen != g_ctx_gen) { h_out = nullptr; owner_gen = g_ctx_gen; }
    if (!h_out) CK(ctx, mfkc_pinned_alloc(ctx, chunk, &h_out));
    for (;;) {
        size_t w = 0;
        CK(ctx, mfkc_emit_next(ctx, (uint8_t *)h_out, chunk, &w));
        if (!w) break;
        if (fwrite(h_out, 1, w, f) != w) die("Can't write %s", out_file.c_str());
    }
    fclose(f);
    std::vector<uint64_t> hist(MFKC_HIST_BINS);
    CK(ctx, mfkc_histogram(ctx, hist.data()));
    if (mfkc_write_stat_file(st_file.c_str(), hist.data()) != MFKC_OK) die("Can't write %s", st_file.c_str());
    uint64_t st[6]; CK(ctx, mfkc_stats(ctx, st));
    report_sample(k, st[0], good, out_file);
    return out_file;
}

// ---- one sample over G GPUs (--gpus G --gpu-mode shard; no reference analogue): the batches of the sample go
// round robin to G contexts, every context stages its k-mers per (owner shard, minimizer bin) in its own HBM, every owner
// counts its bins straight out of all G staging buffers (peer memory), filter + histogram run per shard, and the per-shard
// key-sorted records are merged into one .kmers.bin by mfkc_merge_records.  Byte-identical to the one-GPU file.
std::string count_sample_sharded(std::vector<mfkc_ctx *> &ctxs, int k, int b, const std::vector<std::string> &files, const std::string &name,
                                 const std::string &out_dir, const std::string &st_dir, uint64_t expected_kmers) {
    const uint32_t G = (uint32_t)ctxs.size();
    std::vector<std::unique_ptr<uint8_t[]>> parts(G);         // (not vectors: no zero-fill of gigabytes that are overwritten right away)
    std::vector<uint64_t> n_rec(G, 0), hist(MFKC_HIST_BINS, 0), hs(MFKC_HIST_BINS);
    uint64_t st_sum[6] = {0, 0, 0, 0, 0, 0}, good = 0;
    for (int attempt = 0;; attempt++) {
        uint32_t bins = 0; uint64_t seg_cap = 0, ovf_cap = 0;
        if (mfkc_p2p_bin_geometry(expected_kmers / G + 1, G, k, 0.0, 0.0, &bins, &seg_cap, &ovf_cap) != MFKC_OK) die("bad shard geometry");
        for (auto *c : ctxs) { CK(c, mfkc_reset(c)); CK(c, mfkc_p2p_stage_create_bins(c, bins, seg_cap, ovf_cap)); }
        for (auto *c : ctxs) for (uint32_t r = 0; r < G; r++) CK(c, mfkc_p2p_attach_ctx(c, r, ctxs[r]));
        for (auto *c : ctxs) CK(c, mfkc_p2p_stage_reset(c));
        uint64_t batch = 0;
        mfkc_reader *next = files.empty() ? nullptr : open_reader(files[0]);
        for (size_t i = 0; i < files.size(); i++) {
            mfkc_reader *cur = next;
            next = i + 1 < files.size() ? open_reader(files[i + 1]) : nullptr;
            for_each_batch(ctxs[0], files[i], [&](const uint8_t *bases, const uint64_t *offs, uint32_t n) {
                mfkc_ctx *c = ctxs[batch++ % G];
                CK(c, mfkc_p2p_submit_reads(c, bases, offs, n));
            }, cur);
        }
        std::vector<uint64_t> total(G, 0), per(G);
        for (auto *c : ctxs) { CK(c, mfkc_p2p_counts(c, per.data())); for (uint32_t r = 0; r < G; r++) total[r] += per[r]; }
        bool overflow = false;
        for (uint32_t r = 0; r < G; r++) {
            CK(ctxs[r], mfkc_p2p_drain(ctxs[r], total[r]));
            const int rc = mfkc_flush(ctxs[r]);
            if (rc == MFKC_E_STATE) overflow = true;                       // a staging buffer overflowed: the estimate was too small
            else if (rc != MFKC_OK) die("%s (libmfkc %d)", mfkc_last_error(ctxs[r]), rc);
        }
        if (!overflow) break;
        if (attempt == 3) die("the sample does not fit the staging buffers");
        expected_kmers *= 2;
        warn("sample larger than estimated, counting it again with larger staging buffers");
    }
    for (uint32_t r = 0; r < G; r++) {
        mfkc_ctx *c = ctxs[r];
        uint64_t g = 0;
        CK(c, mfkc_emit_begin(c, b, &g));
        const size_t rs = k > 31 ? 18 : 10, part_bytes = (size_t)g * rs;
        parts[r].reset(new uint8_t[part_bytes + 1]);
        for (size_t pos = 0; pos < part_bytes;) {
            size_t w = 0;
            CK(c, mfkc_emit_next(c, parts[r].get() + pos, part_bytes - pos, &w));
            if (!w) break;
            pos += w;
        }
        n_rec[r] = g; good += g;
        CK(c, mfkc_histogram(c, hs.data()));
        for (int i = 0; i < MFKC_HIST_BINS; i++) hist[i] += hs[i];
        uint64_t st[6]; CK(c, mfkc_stats(c, st));
        for (int i = 0; i < 6; i++) st_sum[i] += st[i];
    }
    mkdirs(out_dir); mkdirs(st_dir);
    const size_t rs = k > 31 ? 18 : 10;
    const std::string out_file = out_dir + "/" + name + (k > 31 ? ".kmers128.bin" : ".kmers.bin"), st_file = st_dir + "/" + name + ".stat.txt";
    const size_t merged_bytes = (size_t)good * rs;
    std::unique_ptr<uint8_t[]> merged(new uint8_t[merged_bytes + 1]);      // first touched by the merge threads, slice by slice
    std::vector<const uint8_t *> pp(G);
    for (uint32_t r = 0; r < G; r++) pp[r] = parts[r].get();
    if (mfkc_merge_records(pp.data(), n_rec.data(), G, (uint32_t)rs, merged.get(), 0) != MFKC_OK) die("record merge failed");
    FILE *f = fopen(out_file.c_str(), "wb");
    if (!f) die("Can't write %s", out_file.c_str());
    if (merged_bytes && fwrite(merged.get(), 1, merged_bytes, f) != merged_bytes) die("Can't write %s", out_file.c_str());
    fclose(f);
    if (mfkc_write_stat_file(st_file.c_str(), hist.data()) != MFKC_OK) die("Can't write %s", st_file.c_str());
    report_sample(k, st_sum[0], good, out_file);
    return out_file;
}

// src/tools/KmersCounterMain.java:103-116
void report_sample(int k, uint64_t size, uint64_t good, const std::string &out_file) {
    info("%s k-mers found, %s (%.1f%%) of them is good (not erroneous)", group_digits(size).c_str(), group_digits(good).c_str(),
         good * 100.0 / size);
    if (size == 0) warn("No k-mers found in reads! Perhaps you reads file is empty or k-mer size is too big");
    else if (good == 0 || good < (uint64_t)(size * 0.03))
        warn("Too few good k-mers were found! Perhaps you should decrease k-mer size or --maximal-bad-frequency value");
    const uint64_t all = k > 31 ? ~0ull : (1ull << (2 * k)) / 2;
    if (size == all) warn("All possible k-mers were found in reads! Perhaps you should increase k-mer size");
    else if (size >= (uint64_t)(all * 0.99)) warn("Almost all possible k-mers were found in reads! Perhaps you should increase k-mer size");
    info("Good k-mers printed to %s", out_file.c_str());
}

mfkc_ctx *make_ctx(int k, const Gpu &g, uint64_t expected_kmers, bool long_kmers = false, int min_seq_len = 0) {
    if (k <= 0) die("The size of k-mer must be at least 1.");               // KmersCounterMain.java:66-69
    if (k > 31 && !long_kmers) die("The size of k-mer must be no more than 31.");   // KmersCounterMain.java:70-73
    if (k > 63) die("The size of k-mer must be no more than 63.");
    mfkc_cfg cfg; memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg; cfg.k = k; cfg.device = g.device; cfg.variant = g.variant;
    cfg.expected_kmers = expected_kmers; cfg.min_seq_len = min_seq_len;
    cfg.n_shards = g.n_shards; cfg.shard_id = g.shard_id;
    mfkc_ctx *ctx = nullptr;
    const int rc = mfkc_create(&cfg, &ctx);
    if (rc != MFKC_OK) die("%s (libmfkc %d)", mfkc_last_error(nullptr), rc);
    g_ctx_gen = ++g_ctx_gen_counter;
    return ctx;
}

uint64_t estimate_bases(const std::vector<std::string> &files) {
    // plain FASTQ: about half of the bytes are bases; gzip: ~4x compression.  Only a sizing hint.
    double b = 0;
    for (const auto &f : files) {
        const double sz = (double)file_size(f);
        const bool gz = ends_with_ci(f, ".gz");
        const bool fq = ends_with_ci(f, ".fastq") || ends_with_ci(f, ".fq") || ends_with_ci(f, ".fastq.gz") || ends_with_ci(f, ".fq.gz");
        b += sz * (gz ? 4.0 : 1.0) * (fq ? 0.5 : 1.0);
    }
    return (uint64_t)b;
}

// ---- kmer-counter-many (src/tools/KmersCounterForManyFilesMain.java:69-120)
int tool_counter(const Args &a, bool many) {
    const int k = parse_int(a, "k", true, 0);
    const int b = parse_int(a, "maximal-bad-frequence", false, 1);
    std::vector<std::string> files = a.many("reads");
    if (files.empty()) die("Missing mandatory parameter --reads");
    const std::string work = a.one("work-dir", "workDir");
    const std::string out_dir = a.one("output-dir", work + "/kmers"), st_dir = a.one("stats-dir", work + "/stats");
    const Gpu g = gpu_opts(a);
    std::vector<std::pair<std::string, std::vector<std::string>>> samples;
    if (many) {
        std::sort(files.begin(), files.end());                             // Arrays.sort(files): path order
        std::vector<std::string> names;
        for (const auto &f : files) names.push_back(reader_name(f));
        for (size_t i = 0; i < files.size();) {
            const bool pair = i + 1 < files.size() &&
                              ((ends_with(names[i], "_r1") && ends_with(names[i + 1], "_r2")) ||
                               (ends_with(names[i], "_R1") && ends_with(names[i + 1], "_R2")));
            if (pair) { samples.push_back({names[i].substr(0, names[i].size() - 3), {files[i], files[i + 1]}}); i += 2; }
            else { samples.push_back({names[i], {files[i]}}); i += 1; }
        }
    } else {
        // KmersCounterMain.getName :122-137
        std::string name;
        const std::string n1 = reader_name(files[0]);
        if (files.size() == 2) {
            const std::string n2 = reader_name(files[1]);
            const bool pair = (ends_with(n1, "_r1") && ends_with(n2, "_r2")) || (ends_with(n1, "_R1") && ends_with(n2, "_R2"));
            name = pair ? n1.substr(0, n1.size() - 3) : n1 + "+";
        } else name = n1 + (files.size() > 1 ? "+" : "");
        samples.push_back({name, files});
    }
    uint64_t biggest = 0;
    for (const auto &s : samples) biggest = std::max(biggest, estimate_bases(s.second));
    // -l / --min-seq-len: the counting call of component-cutter's front half, IOUtils.loadReads(sequences, k, minLen)
    // (src/tools/ComponentCutterMain.java:81-82, src/io/IOUtils.java:761); kmer-counter itself passes 0
    const int min_len = parse_int(a, a.has("min-seq-len") ? "min-seq-len" : "l", false, 0);
    std::vector<std::string> outs(samples.size());
    const bool logical = getenv("MFKC_LOGICAL_GPUS") != nullptr;          // tests: G contexts on one device
    if (g.n_gpus > 1 && (g.mode == "shard" || (g.mode == "auto" && samples.size() < (size_t)g.n_gpus))) {
        // every sample over all GPUs: hash-range sharded, counted out of peer memory, merged (BASELINE config 4)
        if (k > 31) die("--gpu-mode shard serves k <= 31");
        std::vector<mfkc_ctx *> ctxs;
        for (int d = 0; d < g.n_gpus; d++) {
            Gpu gd = g; gd.device = logical ? g.device : g.device + d; gd.n_shards = g.n_gpus; gd.shard_id = d;
            ctxs.push_back(make_ctx(k, gd, biggest / g.n_gpus + 1, false, min_len));
        }
        for (size_t i = 0; i < samples.size(); i++)
            outs[i] = count_sample_sharded(ctxs, k, b, samples[i].second, samples[i].first, out_dir, st_dir, std::max<uint64_t>(estimate_bases(samples[i].second), 1u << 20));
        for (auto *c : ctxs) mfkc_destroy(c);
    } else {
        // batch mode (BASELINE config 3): the samples are independent, so worker threads with one context each take them from
        // a shared counter; no exchange at all.  Two contexts per GPU: while one sample is counted, sorted and written, the
        // next one is already being parsed and copied.  The reference queues one KmersCounterMain per sample
        // (src/tools/KmersCounterForManyFilesMain.java:80-108).
        const int per_gpu = (samples.size() >= 2 * (size_t)g.n_gpus && !getenv("MFKC_ONE_CONTEXT")) ? 2 : 1;
        const int workers = g.n_gpus * per_gpu;
        if (workers == 1) {
            mfkc_ctx *ctx = make_ctx(k, g, biggest, a.has("long-kmers"), min_len);
            for (size_t i = 0; i < samples.size(); i++) outs[i] = count_sample(ctx, k, b, samples[i].second, samples[i].first, out_dir, st_dir);
            mfkc_destroy(ctx);
        } else {
            std::atomic<size_t> next{0};
            std::vector<std::thread> th;
            for (int wk = 0; wk < workers; wk++)
                th.emplace_back([&, wk] {
                    Gpu gd = g; gd.device = logical ? g.device : g.device + wk % g.n_gpus;
                    mfkc_ctx *ctx = make_ctx(k, gd, biggest, a.has("long-kmers"), min_len);
                    for (size_t i; (i = next++) < samples.size();) outs[i] = count_sample(ctx, k, b, samples[i].second, samples[i].first, out_dir, st_dir);
                    mfkc_destroy(ctx);
                });
            for (auto &t : th) t.join();
        }
    }
    for (const auto &o : outs) OUT_VALUE("%s\n", o.c_str());              // "resulting-kmers-files"
    return 0;
}

// ---- ConnectedComponent.loadComponents (src/structures/ConnectedComponent.java:95-122)
void load_components(const std::string &path, std::vector<int64_t> &keys, std::vector<uint64_t> &off) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) die("Can't load components: file not found");
    std::vector<uint8_t> buf((size_t)file_size(path));
    if (fread(buf.data(), 1, buf.size(), f) != buf.size()) die("Can't load components: unknown IOException");
    fclose(f);
    size_t p = 0;
    auto need = [&](size_t n) { if (p + n > buf.size()) die("Can't load components: file corrupted or format mismatch! Do you set a wrong file?"); };
    auto be32 = [&]() { need(4); uint32_t v = 0; for (int i = 0; i < 4; i++) v = (v << 8) | buf[p++]; return (int32_t)v; };
    auto be64 = [&]() { need(8); uint64_t v = 0; for (int i = 0; i < 8; i++) v = (v << 8) | buf[p++]; return (int64_t)v; };
    const int32_t cnt = be32();
    off.assign(1, 0);
    for (int32_t i = 0; i < cnt; i++) {
        const int32_t size = be32();
        (void)be64();                                                       // weight
        for (int32_t j = 0; j < size; j++) keys.push_back(be64());
        off.push_back(keys.size());
    }
}

std::vector<uint8_t> slurp(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) die("Can't load k-mers file %s", path.c_str());
    std::vector<uint8_t> buf((size_t)file_size(path));
    if (fread(buf.data(), 1, buf.size(), f) != buf.size()) die("Can't load k-mers file %s", path.c_str());
    fclose(f);
    if (buf.size() % 10) die("BAD division by work range");                 // src/io/KmersLoadWorker.java:17-19
    return buf;
}

// ---- features-calculator (src/tools/FeaturesCalculatorMain.java:77-236)
int tool_features(const Args &a) {
    const int k = parse_int(a, "k", true, 0);
    if (!a.has("components-file")) die("Missing mandatory parameter --components-file");
    const std::string cm = a.one("components-file");
    const int threshold = parse_int(a, "threshold", false, 0);
    const std::string work = a.one("work-dir", "workDir");
    const Gpu g = gpu_opts(a);
    std::vector<int64_t> keys; std::vector<uint64_t> off;
    load_components(cm, keys, off);
    const uint32_t n_comp = (uint32_t)off.size() - 1;
    info("%s components loaded from %s", group_digits(n_comp).c_str(), cm.c_str());
    if (n_comp == 0) die("No components were found in input files! Can't continue the calculations.");
    const std::string out_dir = work + "/vectors";
    mkdirs(out_dir);
    // --gpus G (records inputs, threshold >= 0): every GPU holds the component set and takes a share of each file's records.
    // A k-mer occurs once per .kmers.bin file, so every key's value lives on exactly one GPU and is 0 on the others:
    // `vec` and `found` add up over the GPUs (0 > threshold is false for threshold >= 0), `cnt` is the same on all.
    // Reads inputs (-i) accumulate instances of the same k-mer and stay on one GPU.
    const bool logical = getenv("MFKC_LOGICAL_GPUS") != nullptr;
    const int G = (g.n_gpus > 1 && threshold >= 0) ? g.n_gpus : 1;
    std::vector<mfkc_ctx *> ctxs;
    for (int d = 0; d < G; d++) { Gpu gd = g; gd.device = logical ? g.device : g.device + d; ctxs.push_back(make_ctx(k, gd, 0)); }
    mfkc_ctx *ctx = ctxs[0];
    if (keys.empty()) keys.push_back(0);
    for (auto *c : ctxs) CK(c, mfkc_fc_load_components(c, keys.data(), off.data(), n_comp));
    const auto sel_files = a.many("selected");
    if (!sel_files.empty()) {
        static const uint8_t none = 0;
        for (auto *c : ctxs) CK(c, mfkc_fc_set_selected(c, &none, 0));      // an (initially empty) active filter
        for (const auto &sf : sel_files) {
            info("Loading file %s...", base_name(sf).c_str());
            auto b = slurp(sf);
            if (!b.empty()) for (auto *c : ctxs) CK(c, mfkc_fc_set_selected(c, b.data(), b.size() / 10));
        }
    }
    std::vector<int64_t> vec(n_comp), vec_g(n_comp); std::vector<uint64_t> found(n_comp), cnt(n_comp), found_g(n_comp), cnt_g(n_comp);
    size_t active = 1;                                                       // contexts that took part in the current input
    auto print_vectors = [&](const std::string &stem, const std::string &shown) {
        CK(ctx, mfkc_fc_features(ctx, threshold, vec.data(), found.data(), cnt.data()));
        for (size_t d = 1; d < active; d++) {
            CK(ctxs[d], mfkc_fc_features(ctxs[d], threshold, vec_g.data(), found_g.data(), cnt_g.data()));
            for (uint32_t i = 0; i < n_comp; i++) { vec[i] = (int64_t)((uint64_t)vec[i] + (uint64_t)vec_g[i]); found[i] += found_g[i]; }
        }
        const std::string vf = out_dir + "/" + stem + ".vec", bf = out_dir + "/" + stem + ".breadth";
        FILE *f = fopen(vf.c_str(), "w"); if (!f) die("Can't write vector to file %s", vf.c_str());
        for (uint32_t i = 0; i < n_comp; i++) fprintf(f, "%lld\n", (long long)vec[i]);
        fclose(f);
        f = fopen(bf.c_str(), "w"); if (!f) die("Can't write vector to file %s", bf.c_str());
        for (uint32_t i = 0; i < n_comp; i++) fprintf(f, "%s\n", java_double((double)found[i] / (double)cnt[i]).c_str());   // 0/0 -> NaN
        fclose(f);
        info("Features for file %s printed to %s", shown.c_str(), vf.c_str());
        info("Components breadth coverage for file %s printed to %s", shown.c_str(), bf.c_str());
        OUT_VALUE("%s\n", vf.c_str());                                     // "features-files"
    };
    for (const auto &rf : a.many("reads")) {                               // :120-134
        active = 1;
        CK(ctx, mfkc_fc_reset_values(ctx));
        for_each_batch(ctx, rf, [&](const uint8_t *bases, const uint64_t *offs, uint32_t n) { CK(ctx, mfkc_fc_add_reads(ctx, bases, offs, n)); });
        print_vectors(reader_name(rf), base_name(rf));
    }
    for (const auto &kf : a.many("kmers")) {                               // :136-163
        active = ctxs.size();
        for (auto *c : ctxs) CK(c, mfkc_fc_reset_values(c));
        info("Loading file %s...", base_name(kf).c_str());
        const auto recs = slurp(kf);
        size_t chunk = 16777200;                                            // src/io/IOUtils.java:30
        if (ctxs.size() > 1) chunk = std::max<size_t>(10, std::min<size_t>(chunk, (recs.size() / 10 + ctxs.size() - 1) / ctxs.size() * 10));   // an even share per GPU
        size_t turn = 0;
        for (size_t p = 0; p < recs.size(); p += chunk, turn++) {           // the chunks go round the GPUs
            const size_t nbytes = std::min(chunk, recs.size() - p);
            mfkc_ctx *c = ctxs[turn % ctxs.size()];
            CK(c, mfkc_fc_add_records(c, recs.data() + p, nbytes / 10));
        }
        std::string stem = base_name(kf);
        if (ends_with_ci(stem, ".kmers.bin")) stem.resize(stem.size() - 10);
        print_vectors(stem, base_name(kf));
    }
    for (auto *c : ctxs) mfkc_destroy(c);
    return 0;
}

// ---- set algebra over .kmers.bin files: kmers-filter, unique-kmers-multi, kmers-samples-counter.
// "-i" / "--k-mers" = input k-mer files (the shared "-i" alias stores them under "reads").
struct KSet {                                                            // one BigLong2ShortHashMap on the device
    mfkc_ctx *ctx; mfkc_kset *h = nullptr;
    explicit KSet(mfkc_ctx *c) : ctx(c) { CK(ctx, mfkc_kset_create(ctx, &h)); }
    ~KSet() { mfkc_kset_destroy(h); }
    KSet(const KSet &) = delete;
    // IOUtils.loadKmers (src/io/IOUtils.java:369-401) incl. its debug lines' data
    void load(const std::vector<std::string> &files, int thr) {
        for (const auto &f : files) {
            info("Loading file %s...", base_name(f).c_str());
            const auto recs = slurp(f);
            const size_t chunk = 16777200;                                  // src/io/IOUtils.java:30
            for (size_t p = 0; p < recs.size(); p += chunk) CK(ctx, mfkc_kset_load_records(h, recs.data() + p, std::min(chunk, recs.size() - p) / 10, thr));
        }
        CK(ctx, mfkc_kset_load_finish(h));
    }
    uint64_t size() const { uint64_t n = 0; mfkc_kset_size(h, &n); return n; }
    // IOUtils.filterAndPrintKmers (src/io/IOUtils.java:101-123) -> number written
    uint64_t print(const KSet *filter, int thr, int fthr, const std::string &out_file) {
        uint64_t good = 0;
        CK(ctx, mfkc_kset_select_begin(h, filter ? filter->h : nullptr, thr, fthr, &good));
        FILE *f = fopen(out_file.c_str(), "wb");
        if (!f) die("Can't write %s", out_file.c_str());
        std::vector<uint8_t> buf(16777200);
        for (;;) {
            size_t w = 0;
            CK(ctx, mfkc_kset_select_next(h, buf.data(), buf.size(), &w));
            if (!w) break;
            if (fwrite(buf.data(), 1, w, f) != w) die("Can't write %s", out_file.c_str());
        }
        fclose(f);
        return good;
    }
};

std::vector<std::string> kmers_inputs(const Args &a) {
    auto v = a.many("reads");
    if (v.empty()) die("Missing mandatory parameter --k-mers");
    return v;
}
void check_k(int k) {
    if (k <= 0) die("The size of k-mer must be at least 1.");
    if (k > 31) die("The size of k-mer must be no more than 31.");
}

// src/tools/KmersFilter.java:76-116
int tool_kmers_filter(const Args &a) {
    const int k = parse_int(a, "k", true, 0); check_k(k);
    const auto inputs = kmers_inputs(a);
    const auto filters = a.many("filter-kmers");
    if (filters.empty()) die("Missing mandatory parameter --filter-kmers");
    const int b = parse_int(a, "maximal-bad-frequence", false, 1), max_thresh = parse_int(a, "max-thresh", false, 0);
    const std::string work = a.one("work-dir", "workDir"), out_dir = a.one("output-dir", work + "/kmers");
    mkdirs(out_dir);
    mfkc_ctx *ctx = make_ctx(k, gpu_opts(a), 0);
    {
        KSet filter_hm(ctx);
        filter_hm.load(filters, b);
        for (const auto &file : inputs) {
            KSet hm(ctx);
            hm.load({file}, b);
            std::string name = base_name(file);
            for (size_t p; (p = name.find(".kmers.bin")) != std::string::npos;) name.erase(p, 10);   // replaceAll(".kmers.bin", "")
            const std::string out_file = out_dir + "/" + name + ".kmers.bin";
            const uint64_t c = hm.print(&filter_hm, b, max_thresh * (int)filters.size(), out_file);
            info("%s k-mers found, %s (%.1f%%) of them survived after filtering", group_digits(hm.size()).c_str(), group_digits(c).c_str(),
                 c * 100.0 / hm.size());
            info("Filtered k-mers printed to %s", out_file.c_str());
            printf("%s\n", out_file.c_str());
        }
    }
    mfkc_destroy(ctx);
    return 0;
}

// src/tools/UniqueKmersMultipleSamplesFinder.java:82-166
int tool_unique_kmers_multi(const Args &a) {
    const int k = parse_int(a, "k", true, 0); check_k(k);
    const auto inputs = kmers_inputs(a);
    const auto filters = a.many("filter-kmers");
    if (filters.empty()) die("Missing mandatory parameter --filter-kmers");
    const int b = parse_int(a, "maximal-bad-frequence", false, 1);
    const int min_s = parse_int(a, "min-samples", false, 1), max_s = parse_int(a, "max-samples", false, 1);
    if (min_s > max_s) die("--min-samples parameter cannot be greater than --max-samples parameter.");
    const std::string work = a.one("work-dir", "workDir"), out_dir = a.one("output-dir", work + "/kmers"), st_dir = a.one("stats-dir", work + "/stats");
    mfkc_ctx *ctx = make_ctx(k, gpu_opts(a), 0);
    {
        KSet hm(ctx), hm_cnt(ctx);
        for (const auto &file : inputs) {
            KSet tmp(ctx);
            tmp.load({file}, b);
            CK(ctx, mfkc_kset_update(hm.h, tmp.h, MFKC_KSET_ADD, b));
            CK(ctx, mfkc_kset_update(hm_cnt.h, tmp.h, MFKC_KSET_INC, b));
        }
        for (const auto &file : filters) {
            KSet filt(ctx);
            filt.load({file}, b);
            CK(ctx, mfkc_kset_update(hm.h, filt.h, MFKC_KSET_ZERO, b));
        }
        mkdirs(out_dir); mkdirs(st_dir);
        for (int i = min_s; i < max_s + 1; i++) {
            const std::string out_file = out_dir + "/filtered_" + std::to_string(i) + ".kmers.bin";
            const uint64_t c = hm.print(&hm_cnt, b, i - 1, out_file);
            info("%s k-mers found, %s (%.1f%%) of them is good (present in one dataset and missing in other)", group_digits(hm.size()).c_str(),
                 group_digits(c).c_str(), c * 100.0 / hm.size());
            info("Good k-mers printed to %s", out_file.c_str());
            printf("%s\n", out_file.c_str());
        }
    }
    mfkc_destroy(ctx);
    return 0;
}

// src/tools/KmersSamplesCounter.java:66-131
int tool_kmers_samples_counter(const Args &a) {
    const int k = parse_int(a, "k", true, 0); check_k(k);
    const auto inputs = kmers_inputs(a);
    const int b = parse_int(a, "maximal-bad-frequence", false, 1);
    const std::string work = a.one("work-dir", "workDir"), out_dir = a.one("output-dir", work + "/kmers"), st_dir = a.one("stats-dir", work + "/stats");
    mkdirs(out_dir); mkdirs(st_dir);
    mfkc_ctx *ctx = make_ctx(k, gpu_opts(a), 0);
    {
        KSet hm(ctx);
        hm.load(inputs, b);
        CK(ctx, mfkc_kset_reset_values(hm.h));
        for (const auto &file : inputs) {
            KSet one(ctx);
            one.load({file}, b);
            CK(ctx, mfkc_kset_update(hm.h, one.h, MFKC_KSET_INC, b));
        }
        const std::string out_file = out_dir + "/n_samples.kmers.bin", st_file = st_dir + "/n_samples.stat.txt";
        const uint64_t c = hm.print(nullptr, 0, 0, out_file);               // IOUtils.printKmers(hm, 0, outFile, stFile)
        static uint64_t hist[MFKC_HIST_BINS];
        CK(ctx, mfkc_kset_histogram(hm.h, hist));
        if (mfkc_write_stat_file(st_file.c_str(), hist) != MFKC_OK) die("Can't write %s", st_file.c_str());
        const uint64_t size = hm.size();
        info("%s k-mers found, %s (%.1f%%) of them is good (not erroneous)", group_digits(size).c_str(), group_digits(c).c_str(), c * 100.0 / size);
        if (size == 0) warn("No k-mers found in reads! Perhaps you reads file is empty or k-mer size is too big");
        else if (c == 0 || c < (uint64_t)(size * 0.03)) warn("Too few good k-mers were found! Perhaps you should decrease k-mer size or --maximal-bad-frequency value");
        const uint64_t all = (1ull << (2 * k)) / 2;
        if (size == all) warn("All possible k-mers were found in reads! Perhaps you should increase k-mer size");
        else if (size >= (uint64_t)(all * 0.99)) warn("Almost all possible k-mers were found in reads! Perhaps you should increase k-mer size");
        info("Good k-mers printed to %s", out_file.c_str());
        printf("%s\n", out_file.c_str());
    }
    mfkc_destroy(ctx);
    return 0;
}

// ---- seq-builder (src/tools/SeqBuilderMain.java:76-168), SURVEY 8f rank 2
int tool_seq_builder(const Args &a) {
    const int k = parse_int(a, "k", true, 0); check_k(k);
    const auto inputs = kmers_inputs(a);
    int b = parse_int(a, "maximal-bad-frequence", false, 1);
    const int len = parse_int(a, a.has("sequence-len") ? "sequence-len" : "l", true, 0);
    const std::string work = a.one("work-dir", "workDir"), out_dir = a.one("output-dir", work + "/sequences");
    mfkc_ctx *ctx = make_ctx(k, gpu_opts(a), 0);
    {
        KSet hm(ctx);
        hm.load(inputs, b);
        static uint64_t hist[MFKC_HIST_BINS];
        CK(ctx, mfkc_kset_histogram(hm.h, hist));
        const int STAT_LEN = 1024;                                         // SeqBuilderMain.java:29
        std::vector<unsigned long long> stat(STAT_LEN, 0);
        unsigned long long total_kmers = 0;
        for (int v = 0; v < MFKC_HIST_BINS; v++) { total_kmers += (unsigned long long)v * hist[v]; stat[std::min(v, STAT_LEN - 1)] += hist[v]; }
        mkdirs(work);
        const std::string dist = work + "/distribution";
        FILE *df = fopen(dist.c_str(), "w"); if (!df) die("Can't write %s", dist.c_str());
        for (int i = 1; i < STAT_LEN; i++) fprintf(df, "%d %llu\n", i, stat[i]);                 // dumpStat :170-176
        fclose(df);
        if (a.has("bottom-cut-percent")) {                                                       // :97-110
            const int bp = parse_int(a, "bottom-cut-percent", true, 0);
            info("Using bottom cut percent = %d", bp);
            const unsigned long long to_cut = total_kmers * (unsigned long long)bp / 100;
            unsigned long long cur = 0;
            for (int i = 0; i < STAT_LEN - 1; i++) {
                if (cur >= to_cut) { b = i; break; }
                cur += (unsigned long long)i * stat[i];
            }
        }
        info("Using maximal bad frequency = %d", b);
        mkdirs(out_dir);
        std::string stem = base_name(inputs[0]);
        if (ends_with(stem, ".kmers.bin")) stem.resize(stem.size() - 10);
        const std::string out_file = out_dir + "/" + stem + (inputs.size() > 1 ? "+" : "") + ".seq.fasta";
        uint64_t ns = 0, nb = 0;
        CK(ctx, mfkc_kset_sequences_begin(hm.h, b, len, &ns, &nb));
        std::vector<uint64_t> off(ns + 1); std::vector<char> bases(nb + 1); std::vector<uint32_t> av(ns + 1), lo(ns + 1), hi(ns + 1);
        CK(ctx, mfkc_kset_sequences_fetch(hm.h, off.data(), bases.data(), av.data(), lo.data(), hi.data()));
        info("%s sequences found", group_digits(ns).c_str());
        if (ns == 0) warn("No sequences were found! Perhaps you should decrease --min-seq-len or --maximal-bad-frequency values");
        FILE *f = fopen(out_file.c_str(), "w"); if (!f) die("Can't write sequences to file");
        for (uint64_t i = 0; i < ns; i++) {                                                      // Sequence.printSequences + FastaDedicatedWriter (70 per line)
            const uint64_t L = off[i + 1] - off[i];
            fprintf(f, ">%llu length=%llu av_weight=%u min_weight=%u max_weight=%u\n", (unsigned long long)(i + 1), (unsigned long long)L, av[i], lo[i], hi[i]);
            for (uint64_t j = 0; j < L; j += 70) { fwrite(bases.data() + off[i] + j, 1, (size_t)std::min<uint64_t>(70, L - j), f); fputc('\n', f); }
        }
        fclose(f);
        info("Sequences printed to %s", out_file.c_str());
        OUT_VALUE("%s\n", out_file.c_str());
    }
    mfkc_destroy(ctx);
    return 0;
}

// ---- seq-builder-many (src/tools/SeqBuilderForManyFilesMain.java:76-92): one seq-builder run per input file, sub-builder
// work directory workDir/sub-builder
int tool_seq_builder_many(const Args &a) {
    if (a.has("maximal-bad-frequence") && a.has("bottom-cut-percent")) die("-b and -bp can not be set both");
    const std::string work = a.one("work-dir", "workDir");
    for (const auto &f : kmers_inputs(a)) {
        Args sub = a;
        sub.tool = "seq-builder";
        sub.opt["reads"] = {f};
        sub.opt["work-dir"] = {work + "/sub-builder"};
        sub.opt["output-dir"] = {a.one("output-dir", work + "/sequences")};
        tool_seq_builder(sub);
    }
    return 0;
}

// ---- component-cutter (src/tools/ComponentCutterMain.java:77-123): IOUtils.loadReads(sequences, k, minLen) on the device,
// the records handed to a device map, ComponentsBuilder.splitStrategy on the device (mfkc_kset_components_*),
// ConnectedComponent.saveComponents (src/structures/ConnectedComponent.java:80-93) + components-stat-<b1>-<b2>.txt
int tool_component_cutter(const Args &a) {
    const int k = parse_int(a, "k", true, 0); check_k(k);
    const int min_len = parse_int(a, a.has("min-seq-len") ? "min-seq-len" : "l", false, 100);
    const int b1 = parse_int(a, "min-component-size", false, 1000), b2 = parse_int(a, "max-component-size", false, 10000);
    const auto files = a.many("reads");
    if (files.empty()) die("Missing mandatory parameter --sequences");
    const std::string work = a.one("work-dir", "workDir");
    const std::string comp_file = a.one("components-file", work + "/components.bin");
    mfkc_ctx *ctx = make_ctx(k, gpu_opts(a), estimate_bases(files), false, min_len);
    for (const auto &f : files)
        for_each_batch(ctx, f, [&](const uint8_t *bases, const uint64_t *offs, uint32_t n) { CK(ctx, mfkc_submit_reads(ctx, bases, offs, n)); });
    CK(ctx, mfkc_flush(ctx));
    uint64_t good = 0;
    CK(ctx, mfkc_emit_begin(ctx, 0, &good));                               // every k-mer of the sequences (count > 0)
    if (good == 0) die("No sequences were found in input files! The following steps will be useless");
    std::vector<uint8_t> recs((size_t)good * 10);
    for (size_t p = 0; p < recs.size();) {
        size_t w = 0;
        CK(ctx, mfkc_emit_next(ctx, recs.data() + p, recs.size() - p, &w));
        if (!w) break;
        p += w;
    }
    info("Searching for components...");
    uint64_t nc = 0, nk = 0;
    std::vector<uint64_t> off; std::vector<int64_t> keys, weight; std::vector<int32_t> thr;
    {
        KSet hm(ctx);
        const size_t chunk = 16777200;
        for (size_t p = 0; p < recs.size(); p += chunk) CK(ctx, mfkc_kset_load_records(hm.h, recs.data() + p, std::min(chunk, recs.size() - p) / 10, 0));
        CK(ctx, mfkc_kset_load_finish(hm.h));
        CK(ctx, mfkc_kset_components_begin(hm.h, b1, b2, &nc, &nk));
        off.resize(nc + 1); keys.resize(nk + 1); weight.resize(nc + 1); thr.resize(nc + 1);
        CK(ctx, mfkc_kset_components_fetch(hm.h, off.data(), keys.data(), weight.data(), thr.data()));
    }
    mfkc_destroy(ctx);
    mkdirs(work);
    const std::string stat_file = work + "/components-stat-" + std::to_string(b1) + "-" + std::to_string(b2) + ".txt";
    FILE *sf = fopen(stat_file.c_str(), "w"); if (!sf) die("Can't write %s", stat_file.c_str());
    fprintf(sf, "# component.no\tcomponent.size\tcomponent.weight\tusedFreqThreshold\n");        // ComponentsBuilder.java:146-151
    for (uint64_t i = 0; i < nc; i++)
        fprintf(sf, "%llu\t%llu\t%lld\t%d\n", (unsigned long long)(i + 1), (unsigned long long)(off[i + 1] - off[i]), (long long)weight[i], thr[i]);
    fclose(sf);
    info("Total %s components were found", group_digits(nc).c_str());
    if (nc == 0) warn("No components were extracted! Perhaps you should decrease --min-component-size value");
    const size_t slash = comp_file.find_last_of('/');
    if (slash != std::string::npos) mkdirs(comp_file.substr(0, slash));
    FILE *f = fopen(comp_file.c_str(), "wb"); if (!f) die("Can't write %s", comp_file.c_str());
    auto be = [&](uint64_t v, int bytes) { for (int i = bytes - 1; i >= 0; i--) fputc((int)((v >> (8 * i)) & 0xFF), f); };
    be((uint32_t)nc, 4);
    for (uint64_t i = 0; i < nc; i++) {
        be((uint32_t)(off[i + 1] - off[i]), 4);
        be((uint64_t)weight[i], 8);
        for (uint64_t j = off[i]; j < off[i + 1]; j++) be((uint64_t)keys[j], 8);
    }
    fclose(f);
    info("Components saved to %s", comp_file.c_str());
    OUT_VALUE("%s\n", comp_file.c_str());                                   // "components-file"
    return 0;
}

// ---- java.util.Formatter for a double: "%.Nf" (FormattedFloatingDecimal: the shortest round-trip digits, rounded HALF_UP
// to N places) and "%s" (Double.toString)
std::string java_format(const std::string &fmt, double d) {
    if (fmt == "%s") return java_double(d);
    int prec = -1;
    if (fmt.size() >= 4 && fmt[0] == '%' && fmt[1] == '.' && fmt.back() == 'f') {
        prec = 0;
        for (size_t i = 2; i + 1 < fmt.size(); i++) { if (fmt[i] < '0' || fmt[i] > '9') { prec = -1; break; } prec = prec * 10 + (fmt[i] - '0'); }
    } else if (fmt == "%f") prec = 6;
    if (prec < 0 || prec > 300) die("--output-format: only %%.<N>f, %%f and %%s are supported here, not '%s'", fmt.c_str());
    if (std::isnan(d)) return "NaN";
    if (std::isinf(d)) return d > 0 ? "Infinity" : "-Infinity";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(d), std::chars_format::scientific);
    std::string sci(buf, r.ptr);
    const size_t e = sci.find('e');
    std::string digits = sci.substr(0, e);
    int point = atoi(sci.c_str() + e + 1) + 1;                              // value = 0.digits x 10^point
    digits.erase(std::remove(digits.begin(), digits.end(), '.'), digits.end());
    // digits wanted = point + prec (may be <= 0)
    const int keep = point + prec;
    std::string kept;
    if (keep < 0) kept = "";
    else if ((size_t)keep >= digits.size()) kept = digits + std::string((size_t)keep - digits.size(), '0');
    else {
        kept = digits.substr(0, (size_t)keep);
        if (digits[(size_t)keep] >= '5') {                                 // HALF_UP
            int i = keep - 1;
            while (i >= 0 && kept[(size_t)i] == '9') { kept[(size_t)i] = '0'; i--; }
            if (i >= 0) kept[(size_t)i]++;
            else { kept = "1" + kept; point++; }
        }
    }
    if (keep < 0) { kept = ""; }
    // kept holds `point` integer digits (if point > 0) followed by prec fraction digits
    std::string ip, fp;
    if (point > 0) { ip = kept.substr(0, (size_t)point); fp = kept.substr((size_t)point); }
    else { ip = "0"; fp = std::string((size_t)std::min(-point, prec), '0') + kept; }
    if ((int)fp.size() < prec) fp += std::string((size_t)prec - fp.size(), '0');
    fp.resize((size_t)prec);
    return std::string(d < 0 || std::signbit(d) ? "-" : "") + ip + (prec ? "." + fp : "");
}

std::string timestamp() {                                                   // Tool.startTimestamp: yyyy.MM.dd_HH.mm.ss
    static std::string ts;
    if (ts.empty()) { char b[64]; time_t t = time(nullptr); struct tm tmv; localtime_r(&t, &tmv); strftime(b, sizeof b, "%Y.%m.%d_%H.%M.%S", &tmv); ts = b; }
    return ts;
}
std::string with_dt(std::string p) { const size_t i = p.find("$DT"); if (i != std::string::npos) p.replace(i, 3, timestamp()); return p; }
std::string remove_ext_ci(const std::string &s, const std::string &ext) { return ends_with_ci(s, ext) ? s.substr(0, s.size() - ext.size()) : s; }

// DistanceMatrixCalculatorMain.printMatrix (src/tools/DistanceMatrixCalculatorMain.java:91-121)
void print_matrix(const std::vector<std::vector<double>> &m, const std::string &path, const std::vector<std::string> *names, const std::vector<int> *perm,
                  const std::string &fmt) {
    const size_t slash = path.find_last_of('/');
    if (slash != std::string::npos) mkdirs(path.substr(0, slash));
    FILE *f = fopen(path.c_str(), "w"); if (!f) die("Failed to print matrix to %s", path.c_str());
    const size_t n = m.size();
    auto at = [&](size_t i) { return perm ? (size_t)(*perm)[i] : i; };
    if (names) { fputc('#', f); for (size_t i = 0; i < n; i++) fprintf(f, "\t%s", (*names)[at(i)].c_str()); fputc('\n', f); }
    for (size_t i = 0; i < n; i++) {
        if (names) fprintf(f, "%s\t", (*names)[at(i)].c_str());
        for (size_t j = 0; j < n; j++) fprintf(f, "%s%s", j ? "\t" : "", java_format(fmt, m[at(i)][at(j)]).c_str());
        fputc('\n', f);
    }
    fclose(f);
}

// ---- dist-matrix-calculator (src/tools/DistanceMatrixCalculatorMain.java:51-153): Bray-Curtis between the .vec files.
// Host arithmetic on n_samples x n_components doubles -- nothing for a GPU here.
std::string tool_dist_matrix(const Args &a) {
    const auto files = a.many("features");
    if (files.empty()) die("Missing mandatory parameter --features");
    std::vector<std::vector<double>> feats;
    for (const auto &ff : files) {                                          // readVector :123-138
        FILE *f = fopen(ff.c_str(), "r"); if (!f) die("Failed to read features from %s", ff.c_str());
        std::vector<double> v; char line[256];
        while (fgets(line, sizeof line, f)) { if (line[0] != '\n' && line[0] != '\r' && line[0]) v.push_back(strtod(line, nullptr)); }
        fclose(f);
        feats.push_back(v);
    }
    const size_t n = feats.size();
    std::vector<std::vector<double>> m(n, std::vector<double>(n, 0.0));
    for (size_t i = 0; i < n; i++)
        for (size_t j = i + 1; j < n; j++) {                                // brayCurtisDistance :140-153, same summation order
            double sumdiff = 0, sum = 0;
            for (size_t p = 0; p < feats[i].size() && p < feats[j].size(); p++) { sumdiff += std::fabs(feats[i][p] - feats[j][p]); sum += std::fabs(feats[i][p]) + std::fabs(feats[j][p]); }
            m[i][j] = m[j][i] = sumdiff / sum;
        }
    const std::string path = with_dt(a.one("matrix-file", a.one("work-dir", "workDir") + "/dist_matrix_$DT_original_order.txt"));
    std::vector<std::string> names;
    for (const auto &ff : files) names.push_back(remove_ext_ci(base_name(ff), ".vec"));
    print_matrix(m, path, a.has("without-names") ? nullptr : &names, nullptr, a.one("output-format", "%.4f"));
    info("Distance matrix printed to %s", path.c_str());
    OUT_VALUE("%s\n", path.c_str());
    return path;
}

// ---- heatmap-maker (src/tools/HeatMapMakerMain.java:91-166), the part that feeds the pipeline's result: parse the matrix
// file (:169-234), cluster the samples by average linkage (FullHeatMap.clusterObjects, src/algo/FullHeatMap.java:218-289),
// renumber them in leaf order (renumber :321-333) and print the renumbered matrix.  The PNG / SVG drawing is out of scope.
std::string tool_heatmap(const Args &a) {
    const std::string in = a.has("matrix-file") ? a.one("matrix-file") : a.one("reads");
    if (in.empty()) die("Missing mandatory parameter --matrix-file");
    FILE *f = fopen(in.c_str(), "r"); if (!f) die("Can't read matrix file %s", in.c_str());
    std::vector<std::vector<std::string>> rows; char *line = nullptr; size_t cap = 0;
    while (getline(&line, &cap, f) >= 0) {
        std::vector<std::string> cells; std::string cur;
        for (char *c = line; *c && *c != '\n' && *c != '\r'; c++) { if (*c == '\t') { if (!cur.empty()) cells.push_back(cur); cur.clear(); } else cur += *c; }
        if (!cur.empty()) cells.push_back(cur);                            // StringTokenizer: empty tokens vanish
        rows.push_back(cells);
    }
    free(line); fclose(f);
    if (rows.empty()) die("No data to read in matrix file %s", in.c_str());
    const size_t fn = rows[0].size();
    if (fn > rows.size()) die("Can't parse matrix, columns' number > rows' number");
    for (size_t i = 0; i < fn; i++) if (rows[i].size() != fn) die("Can't parse matrix, columns' number is different for different rows");
    for (size_t i = fn; i < rows.size(); i++) if (!rows[i].empty()) die("Can't parse matrix, too much rows");
    const bool with_names = fn && rows[0][0] == "#";
    const size_t n = with_names ? fn - 1 : fn, dx = with_names ? 1 : 0;
    std::vector<std::string> names(n);
    std::vector<std::vector<double>> m(n, std::vector<double>(n));
    for (size_t i = 0; i < n; i++) names[i] = with_names ? rows[0][i + 1] : std::to_string(i + 1) + " library";
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < n; j++) m[i][j] = strtod(rows[i + dx][j + dx].c_str(), nullptr);
    // clusterObjects: repeatedly merge the closest pair (first minimum in (i, j) order), group distance = mean over pairs
    struct Node { int no, left, right; };
    std::vector<Node> nodes; std::vector<int> slot(n);
    for (size_t i = 0; i < n; i++) { nodes.push_back(Node{(int)i, -1, -1}); slot[i] = (int)i; }
    std::vector<std::vector<int>> group(n);
    for (size_t i = 0; i < n; i++) group[i] = {(int)i};
    auto gdist = [&](const std::vector<int> &g1, const std::vector<int> &g2) {
        if (g1.empty() || g2.empty()) return -1.0;
        double sum = 0;
        for (int x : g1) for (int y : g2) sum += m[(size_t)x][(size_t)y];
        return sum / (double)g1.size() / (double)g2.size();
    };
    std::vector<std::vector<double>> dist(n, std::vector<double>(n, 0.0));
    for (size_t i = 0; i < n; i++) for (size_t j = i + 1; j < n; j++) dist[i][j] = dist[j][i] = gdist(group[i], group[j]);
    int root = n ? 0 : -1;
    for (size_t count = n; count > 1; count--) {
        double best = 1.7976931348623157e308; int bi = -1, bj = -1;
        for (size_t i = 0; i < n; i++) for (size_t j = i + 1; j < n; j++)
            if (slot[i] >= 0 && slot[j] >= 0 && dist[i][j] < best) { best = dist[i][j]; bi = (int)i; bj = (int)j; }
        if (bi < 0 || best < 0) die("Internal error. Wrong minDist index.");
        nodes.push_back(Node{-1, slot[(size_t)bi], slot[(size_t)bj]});
        root = (int)nodes.size() - 1;
        slot[(size_t)bi] = root; slot[(size_t)bj] = -1;
        group[(size_t)bi].insert(group[(size_t)bi].end(), group[(size_t)bj].begin(), group[(size_t)bj].end());   // getGroup: left leaves, then right
        group[(size_t)bj].clear();
        for (size_t i = 0; i < n; i++) {
            dist[i][(size_t)bj] = dist[(size_t)bj][i] = -1;
            if ((int)i != bi) dist[i][(size_t)bi] = dist[(size_t)bi][i] = gdist(group[(size_t)bi], group[i]);
        }
    }
    std::vector<int> perm;
    if (root >= 0) {                                                        // renumber: leaves left to right
        std::vector<int> stack{root};
        while (!stack.empty()) {
            const Node nd = nodes[(size_t)stack.back()]; stack.pop_back();
            if (nd.no >= 0) perm.push_back(nd.no);
            else { stack.push_back(nd.right); stack.push_back(nd.left); }
        }
    }
    std::string out = a.has("newMatrix-file") ? with_dt(a.one("newMatrix-file")) : remove_ext_ci(in, ".txt") + "_renumbered.txt";
    if (a.has("without-renumbering")) out = in;
    else {
        print_matrix(m, out, &names, &perm, a.one("output-format", "%.4f"));
        info("Renumbered matrix saved to %s", out.c_str());
    }
    warn("The heat map image is not drawn by this build (only the renumbered matrix is produced)");
    OUT_VALUE("%s\n", out.c_str());
    return out;
}

// ---- matrix-builder (src/tools/DistanceMatrixBuilderMain.java:88-176): the reference's default pipeline
//   kmer-counter-many -> seq-builder-many -> component-cutter -> features-calculator -> dist-matrix-calculator -> heatmap-maker
// with the reference's wiring of parameters and default locations.  --output-format is an extra (the reference fixes "%.4f";
// "%s" prints Double.toString, which is what test_data/meta_test_matrix.txt holds).
int tool_matrix_builder(const Args &a) {
    const auto reads = a.many("reads");
    info("Found %zu libraries to process", reads.size());
    if (reads.empty()) die("No libraries to process!!! Can't continue the calculations.");
    const std::string work = a.one("work-dir", "workDir");
    const std::string k = a.one("k", "31"), b = a.one("maximal-bad-frequence", "1");
    const std::string l = a.has("min-seq-len") ? a.one("min-seq-len") : a.one("l", "100");
    const std::string fmt = a.one("output-format", "%.4f");
    auto base = [&]() { Args s; s.opt["k"] = {k}; s.opt["work-dir"] = {work}; if (a.has("gpu")) s.opt["gpu"] = a.many("gpu"); return s; };
    g_sub_tool = true;
    // sample names as kmer-counter-many derives them (sorted paths, _R1/_R2 pairs)
    std::vector<std::string> files = reads; std::sort(files.begin(), files.end());
    std::vector<std::string> names;
    {
        std::vector<std::string> nm; for (const auto &f : files) nm.push_back(reader_name(f));
        for (size_t i = 0; i < nm.size();) {
            const bool pair = i + 1 < nm.size() && ((ends_with(nm[i], "_r1") && ends_with(nm[i + 1], "_r2")) || (ends_with(nm[i], "_R1") && ends_with(nm[i + 1], "_R2")));
            names.push_back(pair ? nm[i].substr(0, nm[i].size() - 3) : nm[i]); i += pair ? 2 : 1;
        }
    }
    { Args s = base(); s.tool = "kmer-counter-many"; s.opt["reads"] = reads; s.opt["maximal-bad-frequence"] = {b}; tool_counter(s, true); }
    std::vector<std::string> kfiles, sfiles;
    for (const auto &nme : names) { kfiles.push_back(work + "/kmers/" + nme + ".kmers.bin"); sfiles.push_back(work + "/sequences/" + nme + ".seq.fasta"); }
    { Args s = base(); s.tool = "seq-builder-many"; s.opt["reads"] = kfiles; s.opt["maximal-bad-frequence"] = {b}; s.opt["sequence-len"] = {l}; tool_seq_builder_many(s); }
    { Args s = base(); s.tool = "component-cutter"; s.opt["reads"] = sfiles; s.opt["min-seq-len"] = {l}; tool_component_cutter(s); }
    {
        Args s = base(); s.tool = "features-calculator"; s.opt["components-file"] = {work + "/components.bin"};
        if (a.has("use-reads-for-calculating-features")) s.opt["reads"] = reads; else s.opt["kmers"] = kfiles;
        tool_features(s);
    }
    std::vector<std::string> vfiles;
    if (a.has("use-reads-for-calculating-features")) for (const auto &f : reads) vfiles.push_back(work + "/vectors/" + reader_name(f) + ".vec");
    else for (const auto &nme : names) vfiles.push_back(work + "/vectors/" + nme + ".vec");
    std::string orig;
    { Args s = base(); s.opt["features"] = vfiles; s.opt["matrix-file"] = {work + "/matrices/dist_matrix_$DT_original_order.txt"}; s.opt["output-format"] = {fmt}; orig = tool_dist_matrix(s); }
    std::string final_matrix;
    { Args s = base(); s.opt["matrix-file"] = {orig}; s.opt["newMatrix-file"] = {a.one("matrix-file", work + "/matrices/dist_matrix_$DT.txt")}; s.opt["output-format"] = {fmt}; final_matrix = tool_heatmap(s); }
    g_sub_tool = false;
    OUT_VALUE("%s\n", final_matrix.c_str());
    return 0;
}

// ---- gen-reads: synthetic FASTQ / FASTA for tests (BASELINE.md section 4 generator)
int tool_gen(int argc, char **argv) {
    // mfkc_cli gen-reads <out.fastq|out.fa> <n_reads> [sample] [total_genome_bp] [n_genomes]
    if (argc < 4) die("usage: mfkc_cli gen-reads <out.fastq|.fa> <n_reads> [sample] [total_genome_bp] [n_genomes]");
    mfkc_synth_cfg c; mfkc_synth_defaults(&c);
    const std::string out = argv[2];
    const uint64_t n = strtoull(argv[3], nullptr, 10);
    if (argc > 4) c.sample = (uint32_t)atoi(argv[4]);
    if (argc > 5) c.total_genome_bp = strtoull(argv[5], nullptr, 10);
    if (argc > 6) c.n_genomes = (uint32_t)atoi(argv[6]);
    const bool fq = ends_with_ci(out, ".fastq") || ends_with_ci(out, ".fq");
    FILE *f = fopen(out.c_str(), "w"); if (!f) die("Can't write %s", out.c_str());
    std::vector<uint8_t> buf((size_t)std::min<uint64_t>(n, 1u << 16) * c.read_len);
    const std::string qual(c.read_len, 'I');
    for (uint64_t s = 0; s < n; s += 1u << 16) {
        const uint64_t m = std::min<uint64_t>(1u << 16, n - s);
        if (mfkc_synth_reads_host(&c, s, m, buf.data()) != MFKC_OK) die("generator failed");
        for (uint64_t i = 0; i < m; i++) {
            const char *r = (const char *)buf.data() + i * c.read_len;
            if (fq) {
                std::string q = qual;
                q[(s + i) % c.read_len] = '5';                              // a char < 64: forces Sanger detection
                for (uint32_t j = 0; j < c.read_len; j++) if (r[j] == 'N') q[j] = '!';
                fprintf(f, "@read_%llu\n%.*s\n+\n%s\n", (unsigned long long)(s + i), (int)c.read_len, r, q.c_str());
            } else fprintf(f, ">read_%llu\n%.*s\n", (unsigned long long)(s + i), (int)c.read_len, r);
        }
    }
    fclose(f);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc > 1 && !strcmp(argv[1], "gen-reads")) return tool_gen(argc, argv);
    const Args a = parse_args(argc, argv);
    // -p / --available-processors (Tool.java:148-151): the reference sizes its worker pools with it; here it bounds the host
    // threads of the readers (parse workers, gzip decoder threads) unless the MFKC_* variables say otherwise
    if (a.has("available-processors") && !a.many("available-processors").empty()) {
        const int p = parse_int(a, "available-processors", false, 0);
        if (p >= 1) {
            setenv("MFKC_READER_THREADS", std::to_string(p).c_str(), 0);
            setenv("MFKC_INFLATE_THREADS", std::to_string(p >= 2 ? p : 1).c_str(), 0);
        }
    }
    if (a.tool == "kmer-counter-many") return tool_counter(a, true);
    if (a.tool == "kmer-counter") return tool_counter(a, false);
    if (a.tool == "features-calculator") return tool_features(a);
    if (a.tool == "kmers-filter") return tool_kmers_filter(a);
    if (a.tool == "unique-kmers-multi") return tool_unique_kmers_multi(a);
    if (a.tool == "kmers-samples-counter") return tool_kmers_samples_counter(a);
    if (a.tool == "seq-builder") return tool_seq_builder(a);
    if (a.tool == "seq-builder-many") return tool_seq_builder_many(a);
    if (a.tool == "component-cutter") return tool_component_cutter(a);
    if (a.tool == "dist-matrix-calculator") { tool_dist_matrix(a); return 0; }
    if (a.tool == "heatmap-maker") { tool_heatmap(a); return 0; }
    if (a.tool == "matrix-builder") return tool_matrix_builder(a);
    die("Tool '%s' is outside the path this build replaces (matrix-builder, kmer-counter-many, kmer-counter, seq-builder, seq-builder-many, "
        "component-cutter, features-calculator, dist-matrix-calculator, heatmap-maker, kmers-filter, unique-kmers-multi, "
        "kmers-samples-counter)", a.tool.c_str());
}
