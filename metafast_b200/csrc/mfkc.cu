// mfkc.cu -- context management and the C ABI (include/mfkc.h) over the sm_100a kernels.
// Device work only; the host-side parser / writers live in host_io.cpp.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/mfkc.h"
#include "kernels.cuh"
#include "bincount.cuh"
#include "kernels128.cuh"
#include "radix_sort.cuh"
#include "kset.cuh"
#include "components.cuh"
#include "components_host.h"
#include "synth.h"

using namespace mfkc;

// ------------------------------------------------------------------------------------------
// profiling slots
// ------------------------------------------------------------------------------------------
enum ProfSlot {
    P_MARK = 0, P_EXTRACT_COUNT, P_EXTRACT_PARTITION, P_DRAIN, P_EXTRACT_BUCKET, P_COUNT_KEYS, P_REHASH, P_CLEAR, P_HIST, P_COMPACT,
    P_SORT, P_RECORDS, P_RLE, P_FC_BUILD, P_FC_RECORDS, P_FC_READS, P_FC_FEATURES, P_GUPS, P_SYNTH, P_BIN_COUNT, P_DRAIN_HEAVY, P_EXTRACT_KEYS, P_NSLOTS
};
// slot names = the kernels they time (extract_skm: extract_skm_kernel<0|3>, extract_skm_shard: <1|2|4> and owner8;
// drain_skm: drain_skm_kernel / drain_p2p_kernel / drain_regions_kernel of the region-blocked table)
static const char *kProfNames[P_NSLOTS] = {
    "mark_read_ends", "extract_direct", "extract_skm", "drain_skm", "extract_skm_shard", "count_keys", "rehash", "table_clear", "table_hist",
    "table_scan", "radix_sort", "records", "rle", "fc_build", "fc_records", "fc_reads", "fc_features",
    "gups", "synth", "bin_count", "drain_heavy", "extract_keys"};

struct PendingTiming { int slot; cudaEvent_t a, b; };

static constexpr int N_STAGE = 8;

struct Staging {
    uint8_t *d_bases = nullptr; size_t cap_bases = 0;
    uint64_t *d_offsets = nullptr; size_t cap_offsets = 0;
    uint32_t *d_flags = nullptr; size_t cap_flags = 0;      // in u32 words
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_copy = nullptr, ev_done = nullptr;
    unsigned long long *h_snap = nullptr;                    // pinned snapshot of Counters::distinct
    bool pending = false;
    uint64_t kmers_submitted_at_end = 0;                     // cumulative upper bound when this batch was queued
};

struct mfkc_ctx {
    mfkc_cfg cfg{};
    int device = 0, sm_count = 148;
    std::string err;
    Staging st[N_STAGE];               // copy streams + staging buffers: the H2D copies run up to N_STAGE batches ahead
                                       // of the kernels (a drain on the compute stream must not stall the copy engine)
    cudaStream_t compute = nullptr;    // every kernel runs here, in submission order ...
    cudaStream_t aux = nullptr;        // ... except the receive side of the shard exchange, which overlaps the next extraction
    cudaStream_t emit = nullptr;       // ... and the count + sort + records kernels of the bin-local mode: LOWEST priority, so that
                                       // the extraction kernels of another context on the same GPU (next sample arriving over
                                       // PCIe) are scheduled first whenever a CTA slot frees up
    cudaEvent_t ev_order = nullptr;
    cudaEvent_t ev_aux = nullptr; bool aux_pending = false;
    int next_buf = 0;

    // hash variant
    Slot *tab = nullptr; uint64_t cap = 0;
    uint64_t tab_alloc_slots = 0;      // size of the allocation behind `tab` (>= cap: a smaller next sample re-uses it)
    Slot128 *tab128 = nullptr; bool k128 = false;   // 32 <= k <= 63: 128-bit keys, 32-byte slots (same cap / regions)
    bool soa = false;                  // region-blocked super-k-mer path, k <= 31: keys[cap] + counts[cap] in ctx->tab
    uint64_t kmers_since_drain = 0; uint32_t drains_since_clear = 0;
    bool host_fed = false;            // the current sample arrives through mfkc_submit_reads (host buffers)
    uint64_t prev_sample_kmers = 0;   // k-mer instances of the previous sample of this context (drain cadence without hints)
    uint64_t prev_sample_kmers_exact = 0;
    bool tab_clean = true;            // no key has been put into the table since it was last cleared
    uint64_t distinct_ub = 0;          // host-side upper bound of occupied slots
    uint64_t kmers_ub_total = 0;       // cumulative upper bound of submitted k-mer instances
    uint64_t max_table_bytes = 0;
    // region-blocked staging (MFKC_VARIANT_HASH)
    unsigned long long *rb_keys = nullptr; unsigned int *rb_cursor = nullptr;
    uint64_t rb_cap = 0;               // staging capacity in keys
    uint64_t rb_cap_max = 0;           // adaptive growth limit
    uint64_t staged_ub = 0;            // upper bound of keys staged since the last drain
    uint32_t n_regions = 1; int region_shift = 19;
    int place = 0;                     // 1: minimizer placement + super-k-mer staging (default for MFKC_VARIANT_HASH)
    bool smem_drain = false;           // regions of <= 2^13 slots drained in shared memory (drain_smem_kernel)
    unsigned long long *sp_ent = nullptr; unsigned int *sp_cursor = nullptr; uint32_t *sp_failed = nullptr;   // spill buffer of the smem drain
    uint32_t sp_cap = 0, sp_failed_cap = 0;
    cudaEvent_t ev_drain = nullptr; bool drain_pending = false; uint64_t kmers_at_drain = 0;
    // tight bound for the region-blocked variant: exact values at the last synchronisation point
    uint64_t distinct_base = 0, kmers_base = 0, recv_since_base = 0;
    unsigned long long *h_drain_snap = nullptr;

    // bin-local counting (bincount.cuh): the default mode of MFKC_VARIANT_HASH for k <= 31 on one GPU.  A sample is staged
    // as a whole, bin by bin, and counted in shared memory when its results are asked for; the global table only takes the
    // heavy bins.  A sample that outgrows the staging buffer falls back to the region-blocked table for good (mode 0).
    int mode = 0;                      // 0: region-blocked table, 1: bin-local
    bool bins_ok = false;              // the context may use mode 1 (variant, k, no explicit table geometry, unsharded)
    bool sample_open = false;          // a batch was submitted since the last reset
    uint32_t os_n_bins = 0; uint64_t os_seg_cap = 0, os_ovf_cap = 0;
    uint64_t os_kmers_budget = 0, os_kmers_staged = 0;
    double os_R = 0, os_rpk = 0;       // previous sample: k-mer instances per distinct k-mer, records per k-mer instance
    double os_good_frac = 0;           // previous count: selected entries per distinct k-mer
    double os_plan_kmers = 0, os_plan_R = 0;      // what the current bin geometry was planned for
    BinCtl *d_binctl = nullptr, *h_binctl = nullptr;
    HeavyEnt *d_heavy = nullptr; uint32_t heavy_cap = 0;
    bool os_counted = false; uint32_t os_thr = 0;             // results of the last count are valid for threshold os_thr
    unsigned long long *os_keys = nullptr; uint16_t *os_counts = nullptr; uint64_t os_n_good = 0;
    uint64_t os_heavy_bins = 0, os_heavy_recs = 0, os_splits = 0, os_ovf = 0, os_total_recs = 0;      // diagnostics of the last count

    // peer-memory shard exchange
    uint4 *p2p_recs = nullptr; unsigned int *p2p_cursor = nullptr; unsigned long long *p2p_kc = nullptr;
    uint64_t p2p_seg_cap = 0; int p2p_log2 = 0;
    int os_mlen = 12;                  // minimizer length the bins of the current sample derive from
    bool p2p_bins = false;             // staging laid out for the bin-local count: (owner, bin) segments + overflow list
    uint32_t p2p_B = 0; uint64_t p2p_ovf_cap = 0, p2p_kmers_in = 0, p2p_n_cursor = 0;
    const uint4 *p2p_peer_recs[P2P_MAX_PEERS] = {nullptr}; const unsigned int *p2p_peer_cursor[P2P_MAX_PEERS] = {nullptr};
    bool p2p_ipc[P2P_MAX_PEERS] = {false};
    const uint4 **d_peer_recs = nullptr; const unsigned int **d_peer_cursor = nullptr;    // device copies of the pointer tables (128-bit drain)

    // sort variant
    unsigned long long *sv_keys = nullptr; uint64_t sv_cap = 0, sv_ub = 0;
    unsigned long long *svs_keys = nullptr; uint32_t *svs_counts = nullptr; uint64_t svs_n = 0;   // compacted state
    unsigned long long *d_bucket_cursor = nullptr; uint64_t *d_bucket_base = nullptr;             // <= 64 shards
    uint64_t *h_bucket = nullptr;                                                                  // pinned

    Counters *d_ctr = nullptr; Counters *h_ctr = nullptr;
    unsigned long long *d_hist = nullptr; uint64_t *h_hist = nullptr; bool hist_valid = false;
    bool dirty = false;                // work submitted since the last flush

    // emit
    unsigned long long *em_keys = nullptr; uint16_t *em_counts = nullptr; uint64_t em_n = 0;
    uint8_t *em_records = nullptr; uint64_t em_cursor = 0; bool em_valid = false;

    // features-calculator
    FcSlot *fc_tab = nullptr; uint64_t fc_cap = 0;
    uint32_t *fc_bloom = nullptr; uint32_t fc_bmask = 0;     // one-hash Bloom bitmap in front of fc_tab (see fc_accumulate)
    unsigned long long *fc_keys = nullptr; uint64_t *fc_off = nullptr; uint32_t fc_ncomp = 0; uint64_t fc_nkeys = 0;
    Slot *fc_sel = nullptr; uint64_t fc_sel_cap = 0; uint64_t fc_sel_n = 0;
    Counters *d_fc_ctr = nullptr;

    // synth
    mfkc_synth_tables *d_synth = nullptr;

    // timing / profile
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool profiling = false;
    double prof_ms[P_NSLOTS] = {0};
    uint64_t prof_launches[P_NSLOTS] = {0};
    std::vector<PendingTiming> pending;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<void *> pinned;
};

static thread_local std::string g_create_err;

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            char b__[512];                                                                        \
            snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            ctx->err = b__;                                                                       \
            return e__ == cudaErrorMemoryAllocation ? MFKC_E_OOM : MFKC_E_CUDA;                   \
        }                                                                                         \
    } while (0)

#define TRY(expr) do { int r__ = (expr); if (r__ != MFKC_OK) return r__; } while (0)

// Temporaries (sort workspace, emit buffers, RLE outputs) come from the stream-ordered pool of the
// device: cudaMalloc/cudaFree of multi-GB blocks cost 10-1000 ms each and serialise the device.
#define TMP_ALLOC(ptr, bytes) CU_TRY(cudaMallocAsync((void **)&(ptr), (size_t)(bytes) ? (size_t)(bytes) : 1, ctx->compute))
#define TMP_FREE(ptr) do { if (ptr) cudaFreeAsync((void *)(ptr), ctx->compute); } while (0)

static int fail(mfkc_ctx *ctx, int code, const char *msg) { if (ctx) ctx->err = msg; return code; }

static cudaEvent_t get_event(mfkc_ctx *ctx) {
    if (!ctx->ev_pool.empty()) { cudaEvent_t e = ctx->ev_pool.back(); ctx->ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ProfScope {           // counts the launch; brackets it with events when profiling is on
    mfkc_ctx *ctx; int slot; cudaStream_t s; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(mfkc_ctx *c, int sl, cudaStream_t st) : ctx(c), slot(sl), s(st) {
        ctx->prof_launches[slot]++;
        if (ctx->profiling) { a = get_event(ctx); b = get_event(ctx); cudaEventRecord(a, s); }
    }
    ~ProfScope() { if (a) { cudaEventRecord(b, s); ctx->pending.push_back({slot, a, b}); } }
};
static void drain_timings(mfkc_ctx *ctx) {
    for (auto &p : ctx->pending) {
        float ms = 0; cudaEventSynchronize(p.b);
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) ctx->prof_ms[p.slot] += ms;
        ctx->ev_pool.push_back(p.a); ctx->ev_pool.push_back(p.b);
    }
    ctx->pending.clear();
}

// work queued on `later` from here on runs after everything queued on `earlier` so far (temporaries are allocated on the
// compute stream; the low-priority emit stream uses them)
static int stream_after(mfkc_ctx *ctx, cudaStream_t later, cudaStream_t earlier) {
    if (later == earlier) return MFKC_OK;
    CU_TRY(cudaEventRecord(ctx->ev_order, earlier));
    CU_TRY(cudaStreamWaitEvent(later, ctx->ev_order, 0));
    return MFKC_OK;
}

static int grid_for(const mfkc_ctx *ctx, uint64_t work_items, int threads, int blocks_per_sm = 8) {
    uint64_t blocks = (work_items + threads - 1) / threads;
    const uint64_t cap = (uint64_t)ctx->sm_count * blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

static int sync_all(mfkc_ctx *ctx) {
    for (int i = 0; i < N_STAGE; i++) CU_TRY(cudaStreamSynchronize(ctx->st[i].stream));
    CU_TRY(cudaStreamSynchronize(ctx->aux));
    CU_TRY(cudaStreamSynchronize(ctx->emit));
    CU_TRY(cudaStreamSynchronize(ctx->compute));
    for (int i = 0; i < N_STAGE; i++) ctx->st[i].pending = false;
    ctx->drain_pending = false; ctx->aux_pending = false;
    return MFKC_OK;
}

static int read_counters(mfkc_ctx *ctx) {      // requires streams idle
    // (not cudaMemcpy: the legacy default stream is shared by every context of the process)
    CU_TRY(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->compute));
    CU_TRY(cudaStreamSynchronize(ctx->compute));
    return MFKC_OK;
}

// ------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------
extern "C" int mfkc_abi_version(void) { return MFKC_ABI_VERSION; }

extern "C" int mfkc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char *mfkc_last_error(const mfkc_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

// cudaMalloc for the long-lived big blocks; if it fails, give the pool's cached memory back first
static cudaError_t big_alloc(mfkc_ctx *ctx, void **p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) return e;
    cudaGetLastError();
    cudaDeviceSynchronize();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) cudaGetLastError();
    return e;
}

// Round a requested capacity to whole regions: cap = n_regions << shift, n_regions <= MAX_REGIONS.
static void plan_regions(const mfkc_ctx *ctx, uint64_t slots, uint64_t *cap, uint32_t *n_regions, int *shift) {
    // default region: 2^17 slots (2 MiB of 16-byte slots).  Measured on cfg2: the drain time does not depend on the
    // region size between 1 and 32 MiB (it is bound by the L2 load->atomic rate, not by capacity), while more,
    // smaller regions spread the cursor atomics of the staging kernel (26.6 ms vs 31 ms at 8 MiB regions).
    // The single-key staging flavour keeps a shared-memory histogram of the regions and is limited to MAX_REGIONS.
    static const int env_shift = getenv("MFKC_REGION_SHIFT") ? atoi(getenv("MFKC_REGION_SHIFT")) : 0;
    int sh = ctx->cfg.region_shift ? (int)ctx->cfg.region_shift : (env_shift ? env_shift : (ctx->smem_drain ? 12 : 17));
    if (slots < (1ull << sh) && !ctx->cfg.region_shift && !ctx->smem_drain) { sh = 10; while ((1ull << sh) < slots) sh++; }
    uint64_t n = (slots + (1ull << sh) - 1) >> sh;
    const uint64_t max_regions = ctx->place ? (uint64_t)MAX_REGIONS_SKM : (uint64_t)MAX_REGIONS;
    while (n > max_regions) { sh++; n = (slots + (1ull << sh) - 1) >> sh; }
    if (n < 1) n = 1;
    *cap = n << sh; *n_regions = (uint32_t)n; *shift = sh;
}

static size_t slot_bytes(const mfkc_ctx *ctx) { return ctx->k128 ? sizeof(Slot128) : (ctx->soa ? 12 : sizeof(Slot)); }
static TabSoA tab_soa_of(void *base, uint64_t cap) {
    TabSoA t; t.keys = reinterpret_cast<unsigned long long *>(base); t.counts = reinterpret_cast<uint32_t *>(t.keys + cap); t.cap = cap; return t;
}
static TabSoA tab_soa(const mfkc_ctx *ctx) { return tab_soa_of(ctx->tab, ctx->cap); }
static TabSoAOps tab_soa_ops(const mfkc_ctx *ctx) { TabSoAOps o; o.t = tab_soa(ctx); return o; }
// Windowed placement for every minimizer-placed table (see placed_upsert_at): a heavy minimizer (poly-A tails, low-complexity
// sequence) can own far more distinct k-mers than one region has slots; with plain linear probing they form one huge cluster
// running through the following regions and every insert walks it (measured: 700 ms for one drain on the GPU that owns poly-A
// at 8 x 20 M reads).  With the window the surplus goes to uniformly spread secondary positions instead.
static uint32_t table_win(const mfkc_ctx *ctx) {
    static const int env_win = getenv("MFKC_WIN") ? atoi(getenv("MFKC_WIN")) : -1;
    if (!ctx->place || ctx->k128 || ctx->soa) return 0u;
    return env_win >= 0 ? (uint32_t)env_win : SMEM_WIN;
}
static TabAoS tab_aos(const mfkc_ctx *ctx) { TabAoS o; o.tab = ctx->tab; o.cap = ctx->cap; return o; }

// plain_aos: a 16-byte-slot helper table (the --selected set) whatever the layout of the main table
static int table_alloc(mfkc_ctx *ctx, uint64_t slots, Slot **out, bool plain_aos = false) {
    Slot *t = nullptr;
    cudaError_t e = big_alloc(ctx, (void **)&t, slots * (plain_aos ? sizeof(Slot) : slot_bytes(ctx)));
    if (e != cudaSuccess) { ctx->err = "cannot allocate k-mer table"; return MFKC_E_OOM; }
    {
        ProfScope ps(ctx, P_CLEAR, ctx->compute);
        if (ctx->k128) table128_clear_kernel<<<grid_for(ctx, 2 * slots, 256, 16), 256, 0, ctx->compute>>>(reinterpret_cast<Slot128 *>(t), slots);
        else if (ctx->soa && !plain_aos) {
            cudaMemsetAsync(t, 0xFF, slots * 8, ctx->compute);                                         // keys = EMPTY
            cudaMemsetAsync(reinterpret_cast<unsigned long long *>(t) + slots, 0, slots * 4, ctx->compute);   // counts = 0
        } else table_clear_kernel<<<grid_for(ctx, slots, 256, 16), 256, 0, ctx->compute>>>(t, slots);
    }
    CU_TRY(cudaGetLastError());
    *out = t;
    return MFKC_OK;
}
static void table_touched(mfkc_ctx *ctx) { ctx->tab_clean = false; }

extern "C" int mfkc_create(const mfkc_cfg *cfg, mfkc_ctx **out) {
    if (!cfg || !out) { g_create_err = "null argument"; return MFKC_E_BADARG; }
    if (cfg->struct_size != sizeof(mfkc_cfg)) { g_create_err = "mfkc_cfg.struct_size mismatch"; return MFKC_E_BADARG; }
    if (cfg->k <= 0) { g_create_err = "The size of k-mer must be at least 1."; return MFKC_E_BADARG; }        // KmersCounterMain.java:66-69
    if (cfg->k > 63) { g_create_err = "The size of k-mer must be no more than 63 (31 in the reference)."; return MFKC_E_BADARG; }
    if (cfg->k > 31 && cfg->variant != MFKC_VARIANT_HASH && cfg->variant != MFKC_VARIANT_HASH_TABLE) {        // the reference stops at 31 (KmersCounterMain.java:70-73)
        g_create_err = "k > 31 (128-bit keys) is only available with MFKC_VARIANT_HASH"; return MFKC_E_BADARG;
    }
    if (cfg->variant != MFKC_VARIANT_HASH && cfg->variant != MFKC_VARIANT_SORT && cfg->variant != MFKC_VARIANT_HASH_DIRECT &&
        cfg->variant != MFKC_VARIANT_HASH_TABLE) {
        g_create_err = "unknown variant"; return MFKC_E_BADARG;
    }
    if (cfg->region_shift && (cfg->region_shift < 4 || cfg->region_shift > 30)) { g_create_err = "bad region_shift"; return MFKC_E_BADARG; }
    if (cfg->n_shards > 64 || (cfg->n_shards > 1 && (cfg->shard_id < 0 || cfg->shard_id >= cfg->n_shards))) {
        g_create_err = "bad shard configuration"; return MFKC_E_BADARG;
    }
    int n_dev = mfkc_device_count();
    if (n_dev <= 0) { g_create_err = "no CUDA device available (libmfkc has no CPU fallback)"; return MFKC_E_CUDA; }
    if (cfg->device < 0 || cfg->device >= n_dev) { g_create_err = "bad device ordinal"; return MFKC_E_BADARG; }

    mfkc_ctx *ctx = new mfkc_ctx();
    ctx->cfg = *cfg;
    ctx->device = cfg->device;
    const bool force_table = cfg->variant == MFKC_VARIANT_HASH_TABLE;      // the region-blocked table for everything
    if (force_table) ctx->cfg.variant = MFKC_VARIANT_HASH;
    cfg = &ctx->cfg;
    auto bail = [&](int code) { g_create_err = ctx->err; mfkc_destroy(ctx); return code; };
#define CR_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e__); return bail(MFKC_E_CUDA); } } while (0)
    CR_TRY(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CR_TRY(cudaGetDeviceProperties(&prop, ctx->device));
    ctx->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < N_STAGE; i++) {
        CR_TRY(cudaStreamCreateWithFlags(&ctx->st[i].stream, cudaStreamNonBlocking));
        CR_TRY(cudaEventCreateWithFlags(&ctx->st[i].ev_copy, cudaEventDisableTiming));
        CR_TRY(cudaEventCreateWithFlags(&ctx->st[i].ev_done, cudaEventDisableTiming));
        CR_TRY(cudaMallocHost(&ctx->st[i].h_snap, sizeof(unsigned long long)));
    }
    {
        int prio_low = 0, prio_high = 0;
        CR_TRY(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
        CR_TRY(cudaStreamCreateWithPriority(&ctx->compute, cudaStreamNonBlocking, prio_high));
        CR_TRY(cudaStreamCreateWithPriority(&ctx->emit, cudaStreamNonBlocking, prio_low));
        CR_TRY(cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming));
    }
    CR_TRY(cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking));
    CR_TRY(cudaEventCreateWithFlags(&ctx->ev_aux, cudaEventDisableTiming));
    CR_TRY(cudaEventCreateWithFlags(&ctx->ev_drain, cudaEventDisableTiming));
    CR_TRY(cudaMallocHost(&ctx->h_drain_snap, sizeof(unsigned long long)));
    CR_TRY(cudaEventCreate(&ctx->t0));
    CR_TRY(cudaEventCreate(&ctx->t1));
    CR_TRY(cudaMalloc(&ctx->d_ctr, sizeof(Counters)));
    CR_TRY(cudaMemset(ctx->d_ctr, 0, sizeof(Counters)));
    CR_TRY(cudaMalloc(&ctx->d_fc_ctr, sizeof(Counters)));
    CR_TRY(cudaMemset(ctx->d_fc_ctr, 0, sizeof(Counters)));
    CR_TRY(cudaMallocHost(&ctx->h_ctr, sizeof(Counters)));
    CR_TRY(cudaMalloc(&ctx->d_hist, MFKC_HIST_BINS * sizeof(unsigned long long)));
    CR_TRY(cudaMallocHost(&ctx->h_hist, MFKC_HIST_BINS * sizeof(uint64_t)));
    CR_TRY(cudaMalloc(&ctx->d_bucket_cursor, 64 * sizeof(unsigned long long)));
    CR_TRY(cudaMalloc(&ctx->d_bucket_base, 64 * sizeof(uint64_t)));
    CR_TRY(cudaMallocHost(&ctx->h_bucket, 64 * sizeof(uint64_t)));
    CR_TRY(cudaMemset(ctx->d_bucket_cursor, 0, 64 * sizeof(unsigned long long)));
    CR_TRY(cudaMemset(ctx->d_bucket_base, 0, 64 * sizeof(uint64_t)));
    CR_TRY(cudaMalloc(&ctx->d_binctl, sizeof(BinCtl)));
    CR_TRY(cudaMemset(ctx->d_binctl, 0, sizeof(BinCtl)));
    CR_TRY(cudaMallocHost(&ctx->h_binctl, sizeof(BinCtl)));

    {   // random 16-byte slot probes: do not let L2 promote a missing sector to a 64/128-byte DRAM fetch
        const char *g = getenv("MFKC_L2_FETCH");
        const size_t gran = g ? (size_t)atoi(g) : 32;
        if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        cudaGetLastError();
    }
    {   // keep freed temporaries cached in the pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        cudaGetLastError();
    }
    size_t free_b = 0, total_b = 0;
    CR_TRY(cudaMemGetInfo(&free_b, &total_b));
    ctx->max_table_bytes = cfg->max_table_bytes ? cfg->max_table_bytes : (uint64_t)(free_b * 0.8);

    if (cfg->variant == MFKC_VARIANT_HASH) {
        const char *sm = getenv("MFKC_STAGE");            // 0 = stage single keys (hash placement), default = super-k-mers
        ctx->place = (sm && atoi(sm) == 0) ? 0 : 1;
        if (cfg->k > 31) { ctx->place = 1; ctx->k128 = true; }
        // MFKC_SOA=1: experimental split key/count layout (faster in the microbenchmark, slower in the real drain)
        { const char *so = getenv("MFKC_SOA"); ctx->soa = ctx->place == 1 && !ctx->k128 && so && atoi(so) == 1; }
        // MFKC_DRAIN=smem: regions small enough (2^12 slots) for one CTA to drain them in shared memory, windowed
        // placement (drain_smem_kernel).  Measured on cfg2 it does not beat the L2-resident drain yet.
        { const char *dm = getenv("MFKC_DRAIN");
          ctx->smem_drain = ctx->place == 1 && !ctx->k128 && !ctx->soa && dm && !strcmp(dm, "smem") &&
                            (cfg->region_shift == 0 || (int)cfg->region_shift <= SMEM_MAX_SHIFT); }
        // bin-local counting (bincount.cuh) unless the caller pinned the table geometry or asked for the table variant
        static const int env_bins = getenv("MFKC_BINS") ? atoi(getenv("MFKC_BINS")) : 1;
        ctx->bins_ok = env_bins && !force_table && ctx->place == 1 && !ctx->k128 && !ctx->soa && !ctx->smem_drain &&
                       cfg->table_slots == 0 && cfg->staging_bytes == 0 && cfg->region_shift == 0;
        if (ctx->bins_ok)
            CR_TRY(cudaFuncSetAttribute(bin_count_kernel<BC_LOG2S, BC_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)bin_count_smem_bytes<BC_LOG2S, BC_THREADS>()));
    }
    if (cfg->variant != MFKC_VARIANT_SORT) {
        uint64_t slots = cfg->table_slots;
        if (!slots && cfg->expected_distinct) slots = cfg->expected_distinct * 2;      // load 0.5
        if (!slots && ctx->bins_ok) slots = 1ull << 20;                               // only heavy bins reach the table
        if (!slots && cfg->expected_kmers && cfg->variant == MFKC_VARIANT_HASH)
            slots = std::min<uint64_t>((uint64_t)(cfg->expected_kmers / 0.85) + 1024, (uint64_t)(total_b * 0.35) / slot_bytes(ctx));
        if (!slots) slots = 1ull << 22;                                                // 64 MiB, grows on demand
        if (slots < 1024) slots = 1024;
        if (slots * slot_bytes(ctx) > ctx->max_table_bytes) slots = ctx->max_table_bytes / slot_bytes(ctx);
        plan_regions(ctx, slots, &slots, &ctx->n_regions, &ctx->region_shift);
        int r = table_alloc(ctx, slots, &ctx->tab);
        if (r != MFKC_OK) return bail(r);
        ctx->cap = slots; ctx->tab_alloc_slots = slots;
        ctx->tab128 = reinterpret_cast<Slot128 *>(ctx->tab);
        CR_TRY(cudaMalloc(&ctx->rb_cursor, MAX_REGIONS_SKM * sizeof(unsigned int)));
        CR_TRY(cudaMemset(ctx->rb_cursor, 0, MAX_REGIONS_SKM * sizeof(unsigned int)));
        ctx->rb_cap_max = (uint64_t)(total_b * 0.25) / 8;
        if (ctx->smem_drain) {
            ctx->sp_cap = 1u << 22;                                     // 4 M spilled instances per drain (64 MiB)
            CR_TRY(cudaMalloc(&ctx->sp_ent, (size_t)ctx->sp_cap * sizeof(unsigned long long)));
            CR_TRY(cudaMalloc(&ctx->sp_cursor, 2 * sizeof(unsigned int)));
            CR_TRY(cudaMemset(ctx->sp_cursor, 0, 2 * sizeof(unsigned int)));
            CR_TRY(cudaFuncSetAttribute(drain_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((16u << SMEM_MAX_SHIFT) + SMEM_SPILL_MAX * sizeof(unsigned long long))));
        }
    }
    CR_TRY(cudaDeviceSynchronize());
#undef CR_TRY
    *out = ctx;
    return MFKC_OK;
}

static void free_bin_outputs(mfkc_ctx *ctx) {
    TMP_FREE(ctx->os_keys); TMP_FREE(ctx->os_counts);
    ctx->os_keys = nullptr; ctx->os_counts = nullptr;
}
static void free_emit(mfkc_ctx *ctx) {
    TMP_FREE(ctx->em_keys); TMP_FREE(ctx->em_counts); TMP_FREE(ctx->em_records);
    ctx->em_keys = nullptr; ctx->em_counts = nullptr; ctx->em_records = nullptr;
    ctx->em_n = 0; ctx->em_cursor = 0; ctx->em_valid = false;
}

extern "C" void mfkc_destroy(mfkc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto &p : ctx->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < N_STAGE; i++) {
        Staging &s = ctx->st[i];
        cudaFree(s.d_bases); cudaFree(s.d_offsets); cudaFree(s.d_flags);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.ev_copy) cudaEventDestroy(s.ev_copy);
        if (s.ev_done) cudaEventDestroy(s.ev_done);
        if (s.h_snap) cudaFreeHost(s.h_snap);
    }
    free_emit(ctx);
    if (ctx->compute) { cudaStreamSynchronize(ctx->compute); cudaStreamDestroy(ctx->compute); ctx->compute = nullptr; }
    if (ctx->aux) { cudaStreamSynchronize(ctx->aux); cudaStreamDestroy(ctx->aux); ctx->aux = nullptr; }
    if (ctx->emit) { cudaStreamSynchronize(ctx->emit); cudaStreamDestroy(ctx->emit); ctx->emit = nullptr; }
    if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
    if (ctx->ev_aux) cudaEventDestroy(ctx->ev_aux);
    cudaFree(ctx->rb_keys); cudaFree(ctx->rb_cursor);
    cudaFree(ctx->sp_ent); cudaFree(ctx->sp_cursor); cudaFree(ctx->sp_failed);
    for (int i = 0; i < P2P_MAX_PEERS; i++) if (ctx->p2p_ipc[i]) { cudaIpcCloseMemHandle((void *)ctx->p2p_peer_recs[i]); cudaIpcCloseMemHandle((void *)ctx->p2p_peer_cursor[i]); }
    cudaFree(ctx->p2p_recs); cudaFree(ctx->p2p_cursor); cudaFree(ctx->p2p_kc); cudaFree(ctx->d_peer_recs); cudaFree(ctx->d_peer_cursor);
    cudaFree(ctx->tab); cudaFree(ctx->sv_keys); cudaFree(ctx->svs_keys); cudaFree(ctx->svs_counts);
    cudaFree(ctx->d_bucket_cursor); cudaFree(ctx->d_bucket_base);
    if (ctx->h_bucket) cudaFreeHost(ctx->h_bucket);
    cudaFree(ctx->d_ctr); cudaFree(ctx->d_fc_ctr); cudaFree(ctx->d_hist);
    if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
    if (ctx->h_hist) cudaFreeHost(ctx->h_hist);
    cudaFree(ctx->fc_tab); cudaFree(ctx->fc_keys); cudaFree(ctx->fc_off); cudaFree(ctx->fc_sel); cudaFree(ctx->fc_bloom);
    cudaFree(ctx->d_synth);
    if (ctx->ev_drain) cudaEventDestroy(ctx->ev_drain);
    if (ctx->h_drain_snap) cudaFreeHost(ctx->h_drain_snap);
    cudaFree(ctx->d_binctl); cudaFree(ctx->d_heavy);
    if (ctx->h_binctl) cudaFreeHost(ctx->h_binctl);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    for (void *p : ctx->pinned) cudaFreeHost(p);
    delete ctx;
}

// kmer-counter-many runs many samples of one study through the same context: size the next sample's table
// from what the previous one needed (its distinct count at load ~0.4) unless the caller fixed the size.
// A denser table is not only smaller to clear and scan: the region-blocked drain touches fewer sectors per
// k-mer and keeps a higher L2 hit rate.
static int resize_for_next_sample(mfkc_ctx *ctx) {
    if (ctx->cfg.variant == MFKC_VARIANT_SORT || !ctx->tab) return MFKC_OK;
    if (read_counters(ctx) != MFKC_OK) return MFKC_OK;
    if (ctx->h_ctr->kmers) ctx->prev_sample_kmers_exact = ctx->h_ctr->kmers;
    if (ctx->mode == 1 && ctx->os_counted && ctx->h_ctr->kmers) {
        // what the bins of the next sample are planned with: instances per distinct k-mer, records per instance
        const uint64_t distinct = ctx->h_ctr->distinct + ctx->h_ctr->bc_distinct;
        ctx->os_R = (double)ctx->h_ctr->kmers / (double)std::max<uint64_t>(distinct, 1);
        ctx->os_rpk = (double)(ctx->os_total_recs + ctx->os_ovf) / (double)ctx->h_ctr->kmers;
    }
    if (ctx->bins_ok || ctx->cfg.table_slots) return MFKC_OK;          // the table only takes heavy bins: keep it as it is
    static const bool off = getenv("MFKC_NO_RESIZE") != nullptr;
    if (off) return MFKC_OK;
    const uint64_t distinct = ctx->h_ctr->distinct;
    if (distinct < (1u << 20)) return MFKC_OK;
    // load 0.4 for device-resident input; host-fed samples get more room (load 0.29): their drains are asynchronous and
    // the bound of reserve_slots must hold with a distinct-count snapshot that lags one drain behind (see count_batch_device)
    static const double env_factor = getenv("MFKC_RESIZE_FACTOR") ? atof(getenv("MFKC_RESIZE_FACTOR")) : 0.0;
    const double factor = env_factor > 0 ? env_factor : (ctx->host_fed ? 3.5 : 2.5);
    uint64_t want = (uint64_t)((double)distinct * factor);
    if ((double)ctx->cap <= 1.5 * (double)want && (double)ctx->cap >= 0.8 * (double)want) return MFKC_OK;
    uint32_t nr = 1; int sh = 19;
    plan_regions(ctx, want, &want, &nr, &sh);
    if (want * slot_bytes(ctx) > ctx->max_table_bytes) return MFKC_OK;
    if (want <= ctx->tab_alloc_slots) {                   // fits the allocation we hold: no cudaFree / cudaMalloc of several GB
        ctx->cap = want; ctx->n_regions = nr; ctx->region_shift = sh;      // (100-150 ms per sample in batch mode); mfkc_reset clears it
        return MFKC_OK;
    }
    CU_TRY(cudaFree(ctx->tab));
    ctx->tab = nullptr; ctx->tab128 = nullptr; ctx->cap = 0; ctx->tab_alloc_slots = 0;
    Slot *nt = nullptr;
    int r = table_alloc(ctx, want, &nt);
    if (r != MFKC_OK) {                                   // should not happen (smaller than what was just freed)
        plan_regions(ctx, 1ull << 22, &want, &nr, &sh);
        TRY(table_alloc(ctx, want, &nt));
    }
    ctx->tab = nt; ctx->tab128 = reinterpret_cast<Slot128 *>(nt); ctx->cap = want; ctx->n_regions = nr; ctx->region_shift = sh;
    ctx->tab_alloc_slots = want;
    return MFKC_OK;
}

extern "C" int mfkc_reset(mfkc_ctx *ctx) {
    if (!ctx) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(sync_all(ctx));
    TRY(resize_for_next_sample(ctx));
    if (ctx->tab && !ctx->tab_clean) {
        ProfScope ps(ctx, P_CLEAR, ctx->compute);
        if (ctx->k128) table128_clear_kernel<<<grid_for(ctx, 2 * ctx->cap, 256, 16), 256, 0, ctx->compute>>>(ctx->tab128, ctx->cap);
        else if (ctx->soa) {
            CU_TRY(cudaMemsetAsync(ctx->tab, 0xFF, ctx->cap * 8, ctx->compute));
            CU_TRY(cudaMemsetAsync(tab_soa(ctx).counts, 0, ctx->cap * 4, ctx->compute));
        } else table_clear_kernel<<<grid_for(ctx, ctx->cap, 256, 16), 256, 0, ctx->compute>>>(ctx->tab, ctx->cap);
        ctx->tab_clean = true;
    }
    ctx->sample_open = false; ctx->os_counted = false; ctx->os_kmers_staged = 0;
    free_bin_outputs(ctx);
    CU_TRY(cudaMemsetAsync(ctx->d_binctl, 0, sizeof(BinCtl), ctx->compute));
    ctx->kmers_since_drain = 0; ctx->drains_since_clear = 0;
    if (ctx->kmers_ub_total) ctx->prev_sample_kmers = ctx->kmers_ub_total;
    CU_TRY(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(Counters), ctx->compute));
    if (ctx->rb_cursor) CU_TRY(cudaMemsetAsync(ctx->rb_cursor, 0, MAX_REGIONS_SKM * sizeof(unsigned int), ctx->compute));
    // no synchronisation: the clear runs in the shadow of the first host-to-device copy of the next sample
    ctx->staged_ub = 0;
    ctx->distinct_base = ctx->kmers_base = ctx->recv_since_base = 0;
    ctx->distinct_ub = 0; ctx->kmers_ub_total = 0; ctx->sv_ub = 0; ctx->svs_n = 0;
    TMP_FREE(ctx->svs_keys); TMP_FREE(ctx->svs_counts); ctx->svs_keys = nullptr; ctx->svs_counts = nullptr;
    ctx->hist_valid = false; ctx->dirty = false;
    free_emit(ctx);
    return MFKC_OK;
}

extern "C" int mfkc_pinned_alloc(mfkc_ctx *ctx, size_t bytes, void **host_ptr) {
    if (!ctx || !host_ptr) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    void *p = nullptr;
    CU_TRY(cudaMallocHost(&p, bytes ? bytes : 1));
    ctx->pinned.push_back(p);
    *host_ptr = p;
    return MFKC_OK;
}
extern "C" int mfkc_pinned_free(mfkc_ctx *ctx, void *host_ptr) {
    if (!ctx) return MFKC_E_BADARG;
    auto it = std::find(ctx->pinned.begin(), ctx->pinned.end(), host_ptr);
    if (it == ctx->pinned.end()) return fail(ctx, MFKC_E_BADARG, "pointer was not allocated by mfkc_pinned_alloc");
    ctx->pinned.erase(it);
    CU_TRY(cudaFreeHost(host_ptr));
    return MFKC_OK;
}

// ------------------------------------------------------------------------------------------
// table growth
// ------------------------------------------------------------------------------------------
static constexpr double kMaxLoad = 0.60;     // never exceeded: checked against an upper bound before each batch
static constexpr double kGrowLoad = 0.30;    // load right after growing (direct variant)
static constexpr double kHardLoad = 0.92;    // region-blocked variant: distinct + staged may reach this fraction of the table

static void poll_snapshots(mfkc_ctx *ctx) {
    if (ctx->drain_pending && cudaEventQuery(ctx->ev_drain) == cudaSuccess) {
        ctx->drain_pending = false;
        const uint64_t cand = *ctx->h_drain_snap + (ctx->kmers_ub_total - ctx->kmers_at_drain);
        if (cand < ctx->distinct_ub) ctx->distinct_ub = cand;
    }
    if (ctx->cfg.variant == MFKC_VARIANT_HASH) {
        // per-batch snapshots of the EXACT number of k-mer instances extracted so far (the host only
        // knows the loose bound bases - k + 1 per batch, which over-counts by ~25 % on 150 bp reads)
        for (int i = 0; i < N_STAGE; i++) {
            Staging &s = ctx->st[i];
            if (s.pending && cudaEventQuery(s.ev_done) == cudaSuccess) {
                s.pending = false;
                const uint64_t exact = *s.h_snap;
                if (exact >= ctx->kmers_base) {
                    const uint64_t cand = ctx->distinct_base + (exact - ctx->kmers_base) + ctx->recv_since_base +
                                          (ctx->kmers_ub_total - s.kmers_submitted_at_end);
                    if (cand < ctx->distinct_ub) ctx->distinct_ub = cand;
                }
            }
        }
        return;
    }
    if (ctx->cfg.variant != MFKC_VARIANT_HASH_DIRECT) return;
    for (int i = 0; i < N_STAGE; i++) {
        Staging &s = ctx->st[i];
        if (s.pending && cudaEventQuery(s.ev_done) == cudaSuccess) {
            s.pending = false;
            const uint64_t cand = *s.h_snap + (ctx->kmers_ub_total - s.kmers_submitted_at_end);
            if (cand < ctx->distinct_ub) ctx->distinct_ub = cand;
        }
    }
}

static int grow_table(mfkc_ctx *ctx, uint64_t need_slots) {
    uint64_t new_cap = std::max<uint64_t>(ctx->cap * 2, need_slots);
    const uint64_t limit = ctx->max_table_bytes / slot_bytes(ctx);
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t fit = (uint64_t)(free_b * 0.95) / slot_bytes(ctx);       // old table stays alive during the rehash
    if (new_cap > limit) new_cap = limit;
    if (new_cap > fit) new_cap = fit;
    uint32_t nr = 1; int sh = 19;
    plan_regions(ctx, new_cap, &new_cap, &nr, &sh);
    while (new_cap > std::min(limit, fit) && new_cap > (1ull << sh)) new_cap -= 1ull << sh, nr--;
    if (new_cap <= ctx->cap) return fail(ctx, MFKC_E_TABLE_FULL, "k-mer table cannot grow: device memory exhausted");
    Slot *nt = nullptr;
    TRY(table_alloc(ctx, new_cap, &nt));
    {
        ProfScope ps(ctx, P_REHASH, ctx->compute);
        if (ctx->k128) {
            TableGeom128 ng; ng.cap = new_cap; ng.n_regions = nr; ng.region_shift = sh; ng.k = ctx->cfg.k;
            rehash128_kernel<<<grid_for(ctx, ctx->cap, 256, 8), 256, 0, ctx->compute>>>(ctx->tab128, ctx->cap, reinterpret_cast<Slot128 *>(nt), ng);
        } else {
            TableGeom ng; ng.cap = new_cap; ng.n_regions = nr; ng.region_shift = sh; ng.k = ctx->cfg.k; ng.minimizer = ctx->place; ng.win = table_win(ctx);
            if (ctx->soa) rehash_soa_kernel<<<grid_for(ctx, ctx->cap, 256, 8), 256, 0, ctx->compute>>>(tab_soa(ctx), tab_soa_of(nt, new_cap), ng);
            else rehash_kernel<<<grid_for(ctx, ctx->cap, 256, 8), 256, 0, ctx->compute>>>(ctx->tab, ctx->cap, nt, ng);
        }
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(ctx->compute));
    CU_TRY(cudaFree(ctx->tab));
    ctx->tab = nt; ctx->cap = new_cap; ctx->tab_alloc_slots = new_cap; ctx->tab128 = reinterpret_cast<Slot128 *>(nt);
    ctx->n_regions = nr; ctx->region_shift = sh;
    return MFKC_OK;
}

// Guarantee that `add` more upserts cannot push the load above kMaxLoad.
static int drain_regions(mfkc_ctx *ctx);

static int reserve_slots(mfkc_ctx *ctx, uint64_t add) {
    // Direct variant: every submitted k-mer may claim a slot at once -> stay below kMaxLoad.
    // Region-blocked variant: staged keys reach the table only at a drain, and the table stays
    // correct (if slower) up to a load close to 1, so drains are forced only by kHardLoad.
    const bool blocked = ctx->cfg.variant == MFKC_VARIANT_HASH;
    const double limit = blocked ? kHardLoad : kMaxLoad;
    poll_snapshots(ctx);
    if ((double)(ctx->distinct_ub + add) <= limit * (double)ctx->cap) { ctx->distinct_ub += add; return MFKC_OK; }
    TRY(drain_regions(ctx));              // staged keys must be in the table before it is measured / rehashed
    TRY(sync_all(ctx));
    TRY(read_counters(ctx));
    ctx->distinct_ub = ctx->distinct_base = ctx->h_ctr->distinct;
    ctx->kmers_base = ctx->h_ctr->kmers; ctx->recv_since_base = 0;
    // grow when the table is really filling up: after growth there is room for as many new keys again
    const uint64_t room = blocked ? std::max<uint64_t>(add, ctx->distinct_ub) : add;
    if ((double)(ctx->distinct_ub + add) > kMaxLoad * (double)ctx->cap) {
        const uint64_t need = (uint64_t)((double)(ctx->distinct_ub + room) / (blocked ? kMaxLoad : kGrowLoad)) + 1024;
        const int r = grow_table(ctx, need);
        if ((double)(ctx->distinct_ub + add) > kHardLoad * (double)ctx->cap) {
            if (r != MFKC_OK) return r;
            return fail(ctx, MFKC_E_TABLE_FULL, "k-mer table cannot grow enough for this batch");
        }
        if (r != MFKC_OK) ctx->err.clear();       // could not grow as far as wished, but the batch fits
    }
    ctx->distinct_ub += add;
    return MFKC_OK;
}

// ------------------------------------------------------------------------------------------
// ingest
// ------------------------------------------------------------------------------------------
static RegionStage region_stage(const mfkc_ctx *ctx) {
    RegionStage rs;
    rs.keys = ctx->rb_keys; rs.cursor = ctx->rb_cursor;
    rs.n_regions = ctx->n_regions; rs.region_shift = ctx->region_shift;
    uint64_t seg = ctx->rb_cap / ctx->n_regions;
    if (seg > 0x7fffffffull) seg = 0x7fffffffull;
    rs.seg_cap = seg;
    return rs;
}

static TableGeom table_geom(const mfkc_ctx *ctx) {
    TableGeom g;
    g.cap = ctx->cap; g.n_regions = ctx->n_regions; g.region_shift = ctx->region_shift; g.k = ctx->cfg.k; g.minimizer = ctx->place; g.win = table_win(ctx);
    return g;
}
static SkmStage skm_stage(const mfkc_ctx *ctx) {
    SkmStage st{};
    st.recs = reinterpret_cast<uint4 *>(ctx->rb_keys); st.cursor = ctx->rb_cursor;
    st.n_regions = ctx->n_regions; st.region_shift = ctx->region_shift; st.win = table_win(ctx);
    uint64_t seg = (ctx->rb_cap / 2) / ctx->n_regions;
    if (seg > 0x7fffffffull) seg = 0x7fffffffull;
    st.seg_cap = seg;
    return st;
}
// staging demand of `kmers` k-mer instances in 8-byte units.  Super-k-mer records are 16 bytes and
// hold (k - m + 2) / 2 k-mers on average for random sequence, capped at 16 by the thread tile; the
// estimate is deliberately high, and a full segment only costs speed (direct upserts), never results.
static uint64_t stage_units(const mfkc_ctx *ctx, uint64_t kmers) {
    if (!ctx->place) return kmers;
    const int w = ctx->cfg.k - minimizer_len(ctx->cfg.k) + 1;
    const int div = std::max(1, std::min(4, w / 4));
    return (ctx->k128 ? 4 : 2) * (kmers / div + 1);               // 32-byte records for 128-bit keys
}
static SkmStage128 skm_stage128(const mfkc_ctx *ctx) {
    SkmStage128 st;
    st.recs = reinterpret_cast<uint4 *>(ctx->rb_keys); st.cursor = ctx->rb_cursor;
    st.n_regions = ctx->n_regions; st.region_shift = ctx->region_shift;
    uint64_t seg = (ctx->rb_cap / 4) / ctx->n_regions;
    if (seg > 0x7fffffffull) seg = 0x7fffffffull;
    st.seg_cap = seg;
    return st;
}

// phase B: upsert every staged key, region by region (asynchronous on the compute stream)
static int drain_regions(mfkc_ctx *ctx) {
    if (ctx->cfg.variant != MFKC_VARIANT_HASH || ctx->staged_ub == 0 || !ctx->rb_keys) return MFKC_OK;
    if (ctx->aux_pending) { CU_TRY(cudaStreamWaitEvent(ctx->compute, ctx->ev_aux, 0)); ctx->aux_pending = false; }
    const uint64_t per_region = ctx->staged_ub / ctx->n_regions + 1;       // 8-byte units
    {
        ProfScope ps(ctx, P_DRAIN, ctx->compute);
        if (ctx->k128) {
            const uint32_t bpr = (uint32_t)std::min<uint64_t>(592, std::max<uint64_t>(1, per_region / 4 / 512));
            drain_skm128_kernel<<<ctx->n_regions * bpr, 256, 0, ctx->compute>>>(skm_stage128(ctx), bpr, ctx->cfg.k, ctx->tab128, ctx->cap, ctx->d_ctr);
        } else if (ctx->smem_drain && ctx->region_shift <= SMEM_MAX_SHIFT) {
            if (ctx->sp_failed_cap < ctx->n_regions) {
                CU_TRY(cudaStreamSynchronize(ctx->compute));
                cudaFree(ctx->sp_failed); ctx->sp_failed = nullptr;
                CU_TRY(cudaMalloc(&ctx->sp_failed, (size_t)ctx->n_regions * sizeof(uint32_t)));
                ctx->sp_failed_cap = ctx->n_regions;
            }
            SpillBuf sp; sp.keys = ctx->sp_ent; sp.cursor = ctx->sp_cursor; sp.n_dirty = ctx->sp_cursor + 1; sp.dirty = ctx->sp_failed; sp.cap = ctx->sp_cap;
            const size_t smem = ((size_t)16 << ctx->region_shift) + SMEM_SPILL_MAX * sizeof(unsigned long long);
            static const bool dbg = getenv("MFKC_DEBUG_DRAIN") != nullptr;
            cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
            if (dbg) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventRecord(e0, ctx->compute); }
            drain_smem_kernel<<<ctx->n_regions, 256, smem, ctx->compute>>>(skm_stage(ctx), ctx->cfg.k, ctx->tab, sp, ctx->d_ctr);
            if (dbg) cudaEventRecord(e1, ctx->compute);
            drain_fallback_kernel<<<ctx->sm_count * 8, 256, 0, ctx->compute>>>(skm_stage(ctx), ctx->cfg.k, ctx->tab, ctx->cap, sp, ctx->d_ctr);
            if (dbg) {
                cudaEventRecord(e2, ctx->compute); cudaEventSynchronize(e2);
                float ma = 0, mb = 0; cudaEventElapsedTime(&ma, e0, e1); cudaEventElapsedTime(&mb, e1, e2);
                unsigned int h[2] = {0, 0}; cudaMemcpy(h, ctx->sp_cursor, sizeof h, cudaMemcpyDeviceToHost);
                fprintf(stderr, "[drain] regions %u (shift %d) seg_cap %llu  smem %.2f ms  fallback %.2f ms  spilled %u  dirty regions %u\n",
                        ctx->n_regions, ctx->region_shift, (unsigned long long)skm_stage(ctx).seg_cap, ma, mb, h[0], h[1]);
                cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
            }
            cudaMemsetAsync(ctx->sp_cursor, 0, 2 * sizeof(unsigned int), ctx->compute);
        } else if (ctx->place) {
            static const int env_bpr = getenv("MFKC_BPR") ? atoi(getenv("MFKC_BPR")) : 0;
            uint32_t bpr = (uint32_t)std::min<uint64_t>(592, std::max<uint64_t>(1, per_region / 2 / 512));
            if (env_bpr > 0 && per_region / 2 > 256) bpr = (uint32_t)env_bpr;
            if (ctx->soa && ctx->drains_since_clear)       // keep the blind u32 increments from ever wrapping (see TabSoA)
                soa_clamp_kernel<<<grid_for(ctx, ctx->cap, 256, 8), 256, 0, ctx->compute>>>(tab_soa(ctx).counts, ctx->cap);
            if (ctx->soa) drain_skm_kernel<TabSoAOps><<<ctx->n_regions * bpr, 256, 0, ctx->compute>>>(skm_stage(ctx), bpr, ctx->cfg.k, tab_soa_ops(ctx), ctx->d_ctr);
            else drain_skm_kernel<TabAoS><<<ctx->n_regions * bpr, 256, 0, ctx->compute>>>(skm_stage(ctx), bpr, ctx->cfg.k, tab_aos(ctx), ctx->d_ctr);
        } else {
            const uint32_t bpr = (uint32_t)std::min<uint64_t>(592, std::max<uint64_t>(1, per_region / 2048));
            drain_regions_kernel<<<ctx->n_regions * bpr, 256, 0, ctx->compute>>>(region_stage(ctx), bpr, ctx->tab, ctx->cap, ctx->d_ctr);
        }
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemsetAsync(ctx->rb_cursor, 0, MAX_REGIONS_SKM * sizeof(unsigned int), ctx->compute));
    // distinct is exact for everything submitted so far once this point of the stream is reached
    CU_TRY(cudaMemcpyAsync(ctx->h_drain_snap, &ctx->d_ctr->distinct, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->compute));
    CU_TRY(cudaEventRecord(ctx->ev_drain, ctx->compute));
    ctx->drain_pending = true;
    ctx->kmers_at_drain = ctx->kmers_ub_total;
    ctx->staged_ub = 0; ctx->kmers_since_drain = 0; ctx->drains_since_clear++;
    return MFKC_OK;
}

// make room for `add` more staged keys (drains, and grows the adaptive staging buffer)
static int reserve_staging(mfkc_ctx *ctx, uint64_t add_kmers) {
    const uint64_t add = stage_units(ctx, add_kmers);
    const uint64_t hard = 4000000000ull;          // cursors are 32-bit
    if (!ctx->rb_keys) {
        uint64_t want = ctx->cfg.staging_bytes ? ctx->cfg.staging_bytes / 8 : std::max<uint64_t>(4 * add, 1ull << 22);
        if (!ctx->cfg.staging_bytes && ctx->cfg.expected_kmers)
            want = std::max<uint64_t>(want, stage_units(ctx, ctx->cfg.expected_kmers + ctx->cfg.expected_kmers / 50) + (uint64_t)std::max<uint32_t>(ctx->n_regions, 65536u) * 64);
        if (!ctx->cfg.staging_bytes && want > ctx->rb_cap_max) want = std::max<uint64_t>(ctx->rb_cap_max, add);
        if (!ctx->cfg.staging_bytes && want > ctx->rb_cap_max) want = std::max<uint64_t>(ctx->rb_cap_max, add);
        if (want < add) want = add;
        cudaError_t e = big_alloc(ctx, (void **)&ctx->rb_keys, want * 8);
        if (e != cudaSuccess) return fail(ctx, MFKC_E_OOM, "cannot allocate the key staging buffer");
        ctx->rb_cap = want;
    }
    if (ctx->staged_ub + add <= std::min(ctx->rb_cap, hard)) return MFKC_OK;
    TRY(drain_regions(ctx));
    const bool adaptive = !ctx->cfg.staging_bytes;
    if (add > ctx->rb_cap || (adaptive && ctx->rb_cap < ctx->rb_cap_max)) {
        // the buffer filled up: double it (bounded), so later drains sweep the table less often
        uint64_t want = std::max<uint64_t>(ctx->rb_cap * 2, add);
        if (adaptive && want > ctx->rb_cap_max) want = std::max<uint64_t>(ctx->rb_cap_max, add);
        if (want > ctx->rb_cap) {
            TRY(sync_all(ctx));
            cudaFree(ctx->rb_keys); ctx->rb_keys = nullptr; ctx->rb_cap = 0;
            cudaError_t e = big_alloc(ctx, (void **)&ctx->rb_keys, want * 8);
            if (e != cudaSuccess) return fail(ctx, MFKC_E_OOM, "cannot grow the key staging buffer");
            ctx->rb_cap = want;
        }
    }
    return MFKC_OK;
}

static int ensure_staging(mfkc_ctx *ctx, Staging &s, uint64_t n_bases, uint64_t n_offsets, bool need_bases) {
    if (need_bases && n_bases + 64 > s.cap_bases) {
        CU_TRY(cudaStreamSynchronize(s.stream)); CU_TRY(cudaStreamSynchronize(ctx->compute));
        cudaFree(s.d_bases); s.d_bases = nullptr;
        s.cap_bases = (size_t)((n_bases + 64) * 1.25) + 4096;
        CU_TRY(cudaMalloc(&s.d_bases, s.cap_bases));
    }
    if (need_bases && n_offsets > s.cap_offsets) {
        CU_TRY(cudaStreamSynchronize(s.stream)); CU_TRY(cudaStreamSynchronize(ctx->compute));
        cudaFree(s.d_offsets); s.d_offsets = nullptr;
        s.cap_offsets = (size_t)(n_offsets * 1.25) + 1024;
        CU_TRY(cudaMalloc(&s.d_offsets, s.cap_offsets * sizeof(uint64_t)));
    }
    const size_t flag_words = (size_t)((n_bases + 31) / 32) + 4;
    if (flag_words > s.cap_flags) {
        CU_TRY(cudaStreamSynchronize(s.stream)); CU_TRY(cudaStreamSynchronize(ctx->compute));
        cudaFree(s.d_flags); s.d_flags = nullptr;
        s.cap_flags = (size_t)(flag_words * 1.25) + 1024;
        CU_TRY(cudaMalloc(&s.d_flags, s.cap_flags * sizeof(uint32_t)));
    }
    return MFKC_OK;
}

// queue K0 (+ memset) for a device-resident batch on the compute stream
static int launch_mark(mfkc_ctx *ctx, Staging &s, const uint64_t *d_offsets, uint32_t n_reads, uint64_t n_bases,
                       int min_len, int count_stats, Counters *ctr) {
    const size_t flag_words = (size_t)((n_bases + 31) / 32) + 4;
    CU_TRY(cudaMemsetAsync(s.d_flags, 0, flag_words * sizeof(uint32_t), ctx->compute));
    {
        ProfScope ps(ctx, P_MARK, ctx->compute);
        mark_read_ends_kernel<<<grid_for(ctx, n_reads, 256, 8), 256, 0, ctx->compute>>>(
            d_offsets, n_reads, n_bases, ctx->cfg.k, min_len, count_stats, s.d_flags, ctr);
    }
    CU_TRY(cudaGetLastError());
    return MFKC_OK;
}

static int extract_grid(const mfkc_ctx *ctx, uint64_t n_bases) {
    const uint64_t tiles = ((n_bases + 15) / 16 + EX_THREADS - 1) / EX_THREADS;
    return grid_for(ctx, tiles * EX_THREADS, EX_THREADS, 8);
}


// ------------------------------------------------------------------------------------------
// bin-local counting (bincount.cuh): planning, staging, the count pass, the way back to the table
// ------------------------------------------------------------------------------------------
// overflow area behind the segments: [list: cap records][meta: cap x 8 B][pool: 2 cap records in chunks of 32][tags: cap / 16 x 8 B]
struct OvfLayout { uint4 *list; uint2 *meta; uint4 *pool; unsigned long long *tags; uint32_t n_chunks; };
static OvfLayout ovf_layout(uint4 *after_segments, uint64_t cap) {
    OvfLayout o;
    o.list = after_segments; o.meta = reinterpret_cast<uint2 *>(o.list + cap);
    o.pool = reinterpret_cast<uint4 *>(o.meta + cap); o.tags = reinterpret_cast<unsigned long long *>(o.pool + 2 * cap);
    o.n_chunks = (uint32_t)(2 * cap / 32);
    return o;
}
static uint64_t ovf_bytes(uint64_t cap) { return cap * 16 + cap * 8 + 2 * cap * 16 + (2 * cap / 32) * 8; }

static SkmStage bin_stage(const mfkc_ctx *ctx) {
    SkmStage st{};
    if (ctx->p2p_bins) {             // sharded: segment (owner * B + bin) of this rank's own staging buffer, extract mode 4
        const uint64_t G = (uint64_t)std::max(1, ctx->cfg.n_shards);
        st.recs = ctx->p2p_recs; st.cursor = ctx->p2p_cursor; st.seg_cap = ctx->p2p_seg_cap;
        st.n_regions = (uint32_t)G; st.region_shift = 0; st.win = ctx->p2p_B;
        const OvfLayout o = ovf_layout(st.recs + G * ctx->p2p_B * ctx->p2p_seg_cap, ctx->p2p_ovf_cap);
        st.ovf = o.list; st.ovf_meta = o.meta; st.ovf_used = ctx->p2p_cursor + G * ctx->p2p_B; st.ovf_cap = (uint32_t)ctx->p2p_ovf_cap;
        st.mlen = ctx->os_mlen;
        return st;
    }
    st.recs = reinterpret_cast<uint4 *>(ctx->rb_keys); st.cursor = ctx->rb_cursor;
    st.seg_cap = ctx->os_seg_cap; st.n_regions = ctx->os_n_bins; st.region_shift = 0; st.win = 0;
    const OvfLayout o = ovf_layout(st.recs + (uint64_t)ctx->os_n_bins * ctx->os_seg_cap, ctx->os_ovf_cap);
    st.ovf = o.list; st.ovf_meta = o.meta; st.ovf_used = &ctx->d_binctl->ovf_cursor; st.ovf_cap = (uint32_t)ctx->os_ovf_cap;
    st.mlen = ctx->os_mlen;
    return st;
}
static BinSrc bin_src(const mfkc_ctx *ctx) {
    BinSrc b{};
    if (ctx->p2p_bins) {             // this shard's bins in every peer's staging buffer (peer memory over NVLink)
        const uint32_t G = (uint32_t)std::max(1, ctx->cfg.n_shards);
        const uint64_t n_seg_recs = (uint64_t)G * ctx->p2p_B * ctx->p2p_seg_cap;
        for (uint32_t i = 0; i < G; i++) {
            b.recs[i] = ctx->p2p_peer_recs[i]; b.cursor[i] = ctx->p2p_peer_cursor[i];
            const OvfLayout o = ovf_layout(const_cast<uint4 *>(ctx->p2p_peer_recs[i]) + n_seg_recs, ctx->p2p_ovf_cap);
            b.ovf[i] = o.pool; b.ovf_tags[i] = o.tags; b.ovf_chunks = o.n_chunks;
        }
        b.seg_cap = ctx->p2p_seg_cap; b.n_src = G; b.seg0 = (uint32_t)ctx->cfg.shard_id * ctx->p2p_B; b.rot = (uint32_t)ctx->cfg.shard_id;
        return b;
    }
    b.recs[0] = reinterpret_cast<const uint4 *>(ctx->rb_keys); b.cursor[0] = ctx->rb_cursor;
    {
        const OvfLayout o = ovf_layout(reinterpret_cast<uint4 *>(ctx->rb_keys) + (uint64_t)ctx->os_n_bins * ctx->os_seg_cap, ctx->os_ovf_cap);
        b.ovf[0] = o.pool; b.ovf_tags[0] = o.tags; b.ovf_chunks = o.n_chunks;
    }
    b.seg_cap = ctx->os_seg_cap; b.n_src = 1; b.seg0 = 0; b.rot = 0;
    return b;
}

// Decide, at the first batch of a sample, whether the sample is counted bin-locally, and lay the staging buffer out:
// n_bins segments of seg_cap records + the overflow list.  Sizes come from the caller's hints or the previous sample
// of this context; a wrong guess costs speed (more heavy bins, more split passes) or sends the sample to the table
// (bins_to_table), never results.
static int plan_bins(mfkc_ctx *ctx, uint64_t first_batch_kmers) {
    ctx->mode = 0;
    if (!ctx->bins_ok || ctx->cfg.n_shards > 1) return MFKC_OK;         // sharded contexts stage through mfkc_p2p_*
    const int k = ctx->cfg.k;
    uint64_t expect = ctx->cfg.expected_kmers ? ctx->cfg.expected_kmers
                    : (ctx->prev_sample_kmers_exact ? ctx->prev_sample_kmers_exact + ctx->prev_sample_kmers_exact / 8 : 0);
    if (!expect) expect = std::max<uint64_t>(4 * first_batch_kmers, 16ull << 20);
    expect = std::max<uint64_t>(expect, first_batch_kmers);
    const double kmers = (double)expect * 1.03 + 4096.0;
    double R = ctx->os_R > 0 ? ctx->os_R : (ctx->cfg.expected_distinct ? kmers / (double)ctx->cfg.expected_distinct : 3.0);
    R = std::min(64.0, std::max(1.0, R));
    // tuning / test knobs (read per sample): slack of the staging segments, planned load of the shared-memory table,
    // a fixed number of bins, a fixed overflow-list capacity
    const double env_slack = getenv("MFKC_BIN_SLACK") ? atof(getenv("MFKC_BIN_SLACK")) : 0.0;
    const double env_load = getenv("MFKC_BIN_LOAD") ? atof(getenv("MFKC_BIN_LOAD")) : 0.0;
    const long env_nbins = getenv("MFKC_BIN_COUNT") ? atol(getenv("MFKC_BIN_COUNT")) : 0;
    const long env_ovf = getenv("MFKC_BIN_OVF") ? atol(getenv("MFKC_BIN_OVF")) : 0;
    const double slack = env_slack > 0.0 ? env_slack : 1.5;          // what does not fit continues in the overflow pool, at no loss
    const double load = env_load > 0 ? env_load : 0.45;
    const double S = (double)(1u << BC_LOG2S);
    // a sample that looks like the one the current geometry was planned for keeps it (kmer-counter-many over similar samples):
    // no re-allocation, no churn from small changes of R
    if (ctx->os_n_bins && !ctx->p2p_bins && ctx->rb_keys && ctx->os_plan_kmers > 0 && kmers <= ctx->os_plan_kmers * 1.05 && kmers >= ctx->os_plan_kmers * 0.8 &&
        R <= ctx->os_plan_R * 1.2 && R >= ctx->os_plan_R * 0.8 && !getenv("MFKC_BIN_COUNT")) {
        const OvfLayout o = ovf_layout(reinterpret_cast<uint4 *>(ctx->rb_keys) + (uint64_t)ctx->os_n_bins * ctx->os_seg_cap, ctx->os_ovf_cap);
        CU_TRY(cudaMemsetAsync(o.tags, 0xFF, (size_t)o.n_chunks * sizeof(unsigned long long), ctx->compute));
        ctx->os_kmers_budget = (uint64_t)(ctx->os_plan_kmers * 1.2);
        ctx->os_kmers_staged = 0;
        ctx->mode = 1;
        return MFKC_OK;
    }
    uint64_t n_bins = (uint64_t)(kmers / R / (load * S)) + 1;
    n_bins = std::min<uint64_t>(std::max<uint64_t>(n_bins, 16), (uint64_t)MAX_REGIONS_SKM);
    if (env_nbins > 0) n_bins = std::min<uint64_t>((uint64_t)env_nbins, (uint64_t)MAX_REGIONS_SKM);
    const int mlen = getenv("MFKC_BIN_M") ? std::min(k, std::max(4, atoi(getenv("MFKC_BIN_M")))) : bin_minimizer_len(k, n_bins);
    const int w = k - mlen + 1;
    const double rpk = (ctx->os_rpk > 0 && mlen == ctx->os_mlen) ? ctx->os_rpk * 1.05 : std::min(1.0, (1.0 / 16 + 2.0 / (w + 1)) * 1.2);
    const double recs = kmers * rpk;
    uint64_t seg_cap = (uint64_t)(recs / (double)n_bins * slack) + 64;
    if (seg_cap > 0x7fffffffull) return MFKC_OK;
    uint64_t ovf_cap = std::min<uint64_t>(std::max<uint64_t>((uint64_t)(recs * 0.06), 1ull << 16), 0x7fffffffull);
    if (env_ovf > 0) ovf_cap = (uint64_t)env_ovf;
    ovf_cap = (ovf_cap + 31) / 32 * 32;                                     // whole chunks of 32 records
    const uint64_t need_units = 2 * n_bins * seg_cap + ovf_bytes(ovf_cap) / 8 + 2;        // 8-byte units of rb_keys: segments + overflow area
    if (need_units > ctx->rb_cap) {
        // (cudaMemGetInfo only when the buffer has to grow: the call takes 1-15 ms on a busy host, with the GPU idle behind it)
        size_t free_b = 0, total_b = 0;
        CU_TRY(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t have = ctx->rb_cap * 8;
        if (need_units * 8 > (uint64_t)((double)total_b * 0.45) || need_units * 8 - have > (uint64_t)((double)free_b * 0.9))
            return MFKC_OK;                                                 // too big to stage as a whole: region-blocked table
    }
    if (need_units > ctx->rb_cap) {
        TRY(sync_all(ctx));
        cudaFree(ctx->rb_keys); ctx->rb_keys = nullptr; ctx->rb_cap = 0;
        if (big_alloc(ctx, (void **)&ctx->rb_keys, need_units * 8) != cudaSuccess) return fail(ctx, MFKC_E_OOM, "cannot allocate the record staging buffer");
        ctx->rb_cap = need_units;
    }
    if (ctx->heavy_cap < n_bins + 4096) {
        TRY(sync_all(ctx));
        cudaFree(ctx->d_heavy); ctx->d_heavy = nullptr;
        ctx->heavy_cap = (uint32_t)std::min<uint64_t>(n_bins + n_bins / 4 + 4096, 0xffffffffull);
        CU_TRY(cudaMalloc(&ctx->d_heavy, (size_t)ctx->heavy_cap * sizeof(HeavyEnt)));
    }
    ctx->os_n_bins = (uint32_t)n_bins; ctx->os_seg_cap = seg_cap; ctx->os_ovf_cap = ovf_cap; ctx->os_mlen = mlen;
    ctx->os_plan_kmers = kmers; ctx->os_plan_R = R;
    {   // no chunk of the overflow pool is taken
        const OvfLayout o = ovf_layout(reinterpret_cast<uint4 *>(ctx->rb_keys) + n_bins * seg_cap, ovf_cap);
        CU_TRY(cudaMemsetAsync(o.tags, 0xFF, (size_t)o.n_chunks * sizeof(unsigned long long), ctx->compute));
    }
    ctx->os_kmers_budget = (uint64_t)(kmers * 1.2);
    ctx->os_kmers_staged = 0;
    ctx->mode = 1;
    return MFKC_OK;
}

static int grow_table(mfkc_ctx *ctx, uint64_t need_slots);
static int reserve_slots(mfkc_ctx *ctx, uint64_t add);

// this context's overflow list -> chunk pool (before anybody counts its bins; idempotent)
static int place_overflow(mfkc_ctx *ctx) {
    const SkmStage st = bin_stage(ctx);
    const OvfLayout o = ovf_layout(st.ovf, st.ovf_cap);
    ovf_place_kernel<<<ctx->sm_count * 4, 256, 0, ctx->compute>>>(st.ovf, st.ovf_meta, st.ovf_used, st.ovf_cap, o.pool, o.tags, o.n_chunks, ctx->d_ctr);
    CU_TRY(cudaGetLastError());
    return MFKC_OK;
}

// heavy entries [e0, e1) and, with ovf, the overflow list -> global table (placement by TableGeom g)
static int launch_heavy(mfkc_ctx *ctx, uint32_t e0, uint32_t e1, bool ovf, const TableGeom &g) {
    ProfScope ps(ctx, P_DRAIN_HEAVY, ctx->compute);
    table_touched(ctx);
    if (e1 > e0) {
        const uint32_t bpe = 2;
        drain_heavy_kernel<<<(e1 - e0) * bpe, 256, 0, ctx->compute>>>(bin_src(ctx), ctx->d_heavy + e0, e1 - e0, bpe, ctx->cfg.k, ctx->tab, g, ctx->d_ctr);
        CU_TRY(cudaGetLastError());
    }
    (void)ovf;                       // overflow chunks belong to their bins: drain_heavy_kernel reads them with the segment
    return MFKC_OK;
}

// records in the overflow list(s) this context has to look at (an upper bound of its own share when sharded)
static int overflow_records(mfkc_ctx *ctx, uint64_t *n) {
    *n = 0;
    if (ctx->p2p_bins) {
        const uint32_t G = (uint32_t)std::max(1, ctx->cfg.n_shards);
        const uint64_t n_seg = (uint64_t)G * ctx->p2p_B;
        for (uint32_t i = 0; i < G; i++) {
            unsigned int c = 0;
            CU_TRY(cudaMemcpyAsync(&c, ctx->p2p_peer_cursor[i] + n_seg, sizeof c, cudaMemcpyDeviceToHost, ctx->compute));
            CU_TRY(cudaStreamSynchronize(ctx->compute));
            *n += std::min<uint64_t>(c, ctx->p2p_ovf_cap);
        }
        return MFKC_OK;
    }
    unsigned int c = 0;
    CU_TRY(cudaMemcpyAsync(&c, &ctx->d_binctl->ovf_cursor, sizeof c, cudaMemcpyDeviceToHost, ctx->compute));
    CU_TRY(cudaStreamSynchronize(ctx->compute));
    *n = std::min<uint64_t>(c, ctx->os_ovf_cap);
    return MFKC_OK;
}

// The sample leaves the bin-local mode (it outgrew the staging buffer, the overflow list is filling up, or a caller
// needs the table): every staged record is counted into the region-blocked table, chunk by chunk so that the table
// grows on exact distinct counts, and the sample continues in mode 0.
static int bins_to_table(mfkc_ctx *ctx) {
    if (ctx->mode != 1) return MFKC_OK;
    if (!ctx->p2p_bins) TRY(place_overflow(ctx));          // (a sharded context filed its list in mfkc_p2p_counts)
    TRY(sync_all(ctx));
    TRY(read_counters(ctx));
    ctx->mode = 0; ctx->os_counted = false;
    free_bin_outputs(ctx);
    if (!ctx->tab_clean) {
        ProfScope ps(ctx, P_CLEAR, ctx->compute);
        table_clear_kernel<<<grid_for(ctx, ctx->cap, 256, 16), 256, 0, ctx->compute>>>(ctx->tab, ctx->cap);
        ctx->tab_clean = true;
    }
    CU_TRY(cudaMemsetAsync(&ctx->d_ctr->distinct, 0, sizeof(unsigned long long), ctx->compute));
    CU_TRY(cudaMemsetAsync(&ctx->d_ctr->bc_distinct, 0, sizeof(unsigned long long), ctx->compute));
    ctx->distinct_ub = ctx->distinct_base = 0; ctx->kmers_base = ctx->h_ctr->kmers; ctx->recv_since_base = 0;
    const uint64_t staged = ctx->p2p_bins ? ctx->p2p_kmers_in : ctx->h_ctr->kmers;      // exact k-mer instances staged for this context
    if (staged) {
        const uint32_t nb = ctx->os_n_bins;
        heavy_all_bins_kernel<<<grid_for(ctx, nb, 256, 4), 256, 0, ctx->compute>>>(ctx->d_heavy, nb);
        CU_TRY(cudaGetLastError());
        // expected size first (one rehash of an empty table instead of a doubling ladder), exact bounds per chunk after
        const double R = ctx->os_R > 0 ? ctx->os_R : 3.0;
        const uint64_t guess = (uint64_t)((double)staged / R / 0.4);
        if (guess > ctx->cap) { if (grow_table(ctx, guess) != MFKC_OK) ctx->err.clear(); }
        uint64_t chunks = (uint64_t)((double)staged * 1.2 / (0.15 * (double)ctx->cap)) + 1;
        if ((double)staged <= kMaxLoad * (double)ctx->cap) chunks = 1;
        chunks = std::min<uint64_t>(chunks, nb);
        for (uint64_t c = 0; c < chunks; c++) {
            const uint32_t b0 = (uint32_t)(nb * c / chunks), b1 = (uint32_t)(nb * (c + 1) / chunks);
            const uint64_t share = chunks == 1 ? staged : (uint64_t)((double)staged * (double)(b1 - b0) / (double)nb * 1.2) + 1024;
            TRY(reserve_slots(ctx, share));
            ctx->recv_since_base += share;
            TRY(launch_heavy(ctx, b0, b1, false, table_geom(ctx)));
        }
        TRY(sync_all(ctx));
    }
    if (!ctx->p2p_bins) {            // (a sharded context leaves the staging buffers alone: the peers read them too)
        CU_TRY(cudaMemsetAsync(ctx->rb_cursor, 0, MAX_REGIONS_SKM * sizeof(unsigned int), ctx->compute));
        CU_TRY(cudaMemsetAsync(ctx->d_binctl, 0, sizeof(BinCtl), ctx->compute));
    }
    TRY(sync_all(ctx));
    TRY(read_counters(ctx));
    ctx->distinct_ub = ctx->distinct_base = ctx->h_ctr->distinct;
    ctx->kmers_base = ctx->h_ctr->kmers; ctx->recv_since_base = 0;
    ctx->staged_ub = 0; ctx->kmers_since_drain = 0;
    return MFKC_OK;
}

// a table of at least `need_keys / 0.6` slots, empty, for the heavy bins of one count pass (plain hash placement)
static int resid_table_prepare(mfkc_ctx *ctx, uint64_t need_keys) {
    uint64_t want = (uint64_t)((double)need_keys / kMaxLoad) + 1024;
    if (want > ctx->cap) {
        if (want * sizeof(Slot) > ctx->max_table_bytes) return fail(ctx, MFKC_E_TABLE_FULL, "the heavy bins do not fit the table");
        uint32_t nr = 1; int sh = 17;
        plan_regions(ctx, want, &want, &nr, &sh);
        CU_TRY(cudaStreamSynchronize(ctx->compute));
        CU_TRY(cudaFree(ctx->tab));
        ctx->tab = nullptr; ctx->tab128 = nullptr; ctx->cap = 0; ctx->tab_alloc_slots = 0;
        Slot *nt = nullptr;
        TRY(table_alloc(ctx, want, &nt));
        ctx->tab = nt; ctx->tab128 = reinterpret_cast<Slot128 *>(nt); ctx->cap = want; ctx->tab_alloc_slots = want;
        ctx->n_regions = nr; ctx->region_shift = sh;
        ctx->tab_clean = true;
    } else if (!ctx->tab_clean) {
        ProfScope ps(ctx, P_CLEAR, ctx->compute);
        table_clear_kernel<<<grid_for(ctx, ctx->cap, 256, 16), 256, 0, ctx->compute>>>(ctx->tab, ctx->cap);
        ctx->tab_clean = true;
    }
    return MFKC_OK;
}

// entry points that put keys into the table themselves (receive sides of the shard exchange)
static int ensure_table_mode(mfkc_ctx *ctx) {
    if (ctx->mode == 1) TRY(bins_to_table(ctx));
    ctx->mode = 0; ctx->sample_open = true;
    table_touched(ctx);
    return MFKC_OK;
}

static constexpr uint32_t kNoOutput = 0xFFFFFFFFu;

// The count pass of the bin-local mode: histogram, distinct count and (thr != kNoOutput) the compacted entries with
// count > thr in ctx->os_keys / os_counts (unsorted).  The staged records stay where they are, so the pass can be
// repeated for another threshold.  Returns 1 when the sample had to fall back to the table (ctx->mode is 0 then).
static int bins_count(mfkc_ctx *ctx, uint32_t thr) {
    const bool want_out = thr != kNoOutput;
    if (ctx->os_counted && (!want_out || (ctx->os_thr == thr && (ctx->os_keys || ctx->os_n_good == 0)))) return MFKC_OK;
    static const bool low = getenv("MFKC_EMIT_PRIORITY") == nullptr || atoi(getenv("MFKC_EMIT_PRIORITY")) != 0;
    cudaStream_t st = low ? ctx->emit : ctx->compute;
    if (!ctx->p2p_bins) TRY(place_overflow(ctx));          // (a sharded context filed its list in mfkc_p2p_counts)
    TRY(sync_all(ctx));
    TRY(read_counters(ctx));
    if (ctx->h_ctr->overflow) return fail(ctx, MFKC_E_STATE, "internal key buffer overflow");
    free_bin_outputs(ctx);
    const uint64_t kmers = ctx->p2p_bins ? ctx->p2p_kmers_in : ctx->h_ctr->kmers;
    uint64_t out_cap = 0;
    if (want_out) {
        const double R = ctx->os_R > 0 ? ctx->os_R : 3.0;
        out_cap = ctx->os_good_frac > 0 ? (uint64_t)((double)kmers / R * ctx->os_good_frac * 1.3) + (1u << 16) : kmers / 6 + (1u << 16);
        out_cap = std::min<uint64_t>(out_cap, kmers);
    }
    static const int env_tma = getenv("MFKC_BIN_TMA") ? atoi(getenv("MFKC_BIN_TMA")) : 1;
    uint64_t good = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        unsigned long long *ok = nullptr; uint16_t *oc = nullptr;
        if (out_cap) { TMP_ALLOC(ok, out_cap * 8); TMP_ALLOC(oc, out_cap * 2); }
        TRY(stream_after(ctx, st, ctx->compute));
        CU_TRY(cudaMemsetAsync(ctx->d_hist, 0, MFKC_HIST_BINS * sizeof(unsigned long long), st));
        CU_TRY(cudaMemsetAsync(&ctx->d_ctr->n_good, 0, 2 * sizeof(unsigned long long), st));            // n_good, bc_distinct
        CU_TRY(cudaMemsetAsync(&ctx->d_ctr->distinct, 0, sizeof(unsigned long long), st));
        CU_TRY(cudaMemsetAsync(ctx->d_binctl, 0, offsetof(BinCtl, ovf_cursor), st));
        BinCountArgs a{};
        a.src = bin_src(ctx); a.n_bins = ctx->os_n_bins; a.k = ctx->cfg.k; a.thr = want_out ? thr : 0x7FFFFFFFu;
        a.limit = (uint32_t)(0.8 * (double)(1u << BC_LOG2S)); a.max_recs = 0xFFFFFFFFu; a.use_tma = env_tma;
        // CTAs retire after a few dozen bins (~1 ms): a second context's extraction kernels (next sample arriving over PCIe
        // while this one is counted) get SMs at every turnover instead of waiting for the whole count
        static const int env_bpc = getenv("MFKC_BIN_PER_CTA") ? atoi(getenv("MFKC_BIN_PER_CTA")) : 0;
        a.bins_per_cta = env_bpc > 0 ? (uint32_t)env_bpc : 32u;
        a.out_keys = ok; a.out_counts = oc; a.out_cap = out_cap; a.hist = ctx->d_hist; a.ctr = ctx->d_ctr; a.ctl = ctx->d_binctl;
        a.heavy = ctx->d_heavy; a.heavy_cap = ctx->heavy_cap;
        if (kmers) {
            ProfScope ps(ctx, P_BIN_COUNT, st);
            const int grid = (int)((ctx->os_n_bins + a.bins_per_cta - 1) / a.bins_per_cta);
            bin_count_kernel<BC_LOG2S, BC_THREADS><<<grid, BC_THREADS, bin_count_smem_bytes<BC_LOG2S, BC_THREADS>(), st>>>(a);
        }
        CU_TRY(cudaGetLastError());
        // one read-back for the common case (no heavy bins): control block, histogram and counters together
        CU_TRY(cudaMemcpyAsync(ctx->h_binctl, ctx->d_binctl, sizeof(BinCtl), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(ctx->h_hist, ctx->d_hist, MFKC_HIST_BINS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        const BinCtl c = *ctx->h_binctl;
        uint64_t n_ovf = std::min<uint64_t>(c.ovf_cursor, ctx->os_ovf_cap);
        if (ctx->p2p_bins) TRY(overflow_records(ctx, &n_ovf));
        ctx->os_heavy_bins = c.n_heavy; ctx->os_heavy_recs = c.heavy_recs; ctx->os_splits = c.n_split; ctx->os_ovf = n_ovf; ctx->os_total_recs = c.total_recs;
        if (c.heavy_overflow) {                                  // more heavy entries than the list holds: the table takes the sample
            TMP_FREE(ok); TMP_FREE(oc);
            TRY(bins_to_table(ctx));
            return 1;
        }
        if (c.n_heavy) {
            const int r = resid_table_prepare(ctx, 16 * c.heavy_recs);
            if (r != MFKC_OK) { TMP_FREE(ok); TMP_FREE(oc); ctx->err.clear(); TRY(bins_to_table(ctx)); return 1; }
            TableGeom g = table_geom(ctx); g.minimizer = 0; g.win = 0;
            TRY(launch_heavy(ctx, 0, c.n_heavy, false, g));               // (on the compute stream, like the table clear before)
            TRY(stream_after(ctx, st, ctx->compute));
            {
                ProfScope ps(ctx, P_COMPACT, st);
                table_scan_kernel<false><<<grid_for(ctx, (ctx->cap + 7) / 8, 256, 8), 256, 0, st>>>(
                    ctx->tab, TabSoA{nullptr, nullptr, 0}, ctx->cap, a.thr, ctx->d_hist, ok, oc, out_cap, ctx->d_ctr);
            }
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaMemcpyAsync(ctx->h_hist, ctx->d_hist, MFKC_HIST_BINS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
        }
        good = ctx->h_ctr->n_good;
        if (good <= out_cap || !want_out) { ctx->os_keys = ok; ctx->os_counts = oc; break; }
        TMP_FREE(ok); TMP_FREE(oc);
        if (attempt == 1) return fail(ctx, MFKC_E_STATE, "internal: bin count output overflow");
        out_cap = good;                                           // exact now
    }
    if (!want_out) good = 0;
    ctx->os_n_good = good; ctx->os_thr = thr; ctx->os_counted = true; ctx->hist_valid = true;
    const uint64_t distinct = ctx->h_ctr->distinct + ctx->h_ctr->bc_distinct;
    if (want_out && distinct) ctx->os_good_frac = (double)good / (double)distinct;
    ctx->distinct_ub = ctx->distinct_base = ctx->h_ctr->distinct;
    return MFKC_OK;
}

static int sort_variant_reserve(mfkc_ctx *ctx, uint64_t add);

// count a device-resident batch (bases/offsets already on the device, visible to the compute stream)
static int count_batch_device(mfkc_ctx *ctx, Staging &s, const uint8_t *d_bases, const uint64_t *d_offsets,
                              uint32_t n_reads, uint64_t n_bases, bool host_fed = false) {
    const int k = ctx->cfg.k;
    const uint64_t kmers_ub = n_bases >= (uint64_t)k ? n_bases - k + 1 : 0;
    if (!ctx->sample_open) {
        ctx->sample_open = true;
        if (ctx->cfg.variant == MFKC_VARIANT_HASH) TRY(plan_bins(ctx, kmers_ub));
    }
    ctx->os_counted = false;
    if (ctx->mode == 1) {
        // bin-local mode: stage only.  The sample stays in this mode while it fits the plan and the overflow list
        // (watched through asynchronous snapshots of its cursor) stays far from full.
        uint64_t kmers_est = n_bases >= (uint64_t)n_reads * (uint64_t)(k - 1) ? n_bases - (uint64_t)n_reads * (uint64_t)(k - 1) : 0;
        uint64_t ovf_seen = 0;
        for (int i = 0; i < N_STAGE; i++) {
            Staging &q = ctx->st[i];
            if (q.pending && cudaEventQuery(q.ev_done) == cudaSuccess) { q.pending = false; ovf_seen = std::max<uint64_t>(ovf_seen, *q.h_snap); }
        }
        if (ctx->os_kmers_staged + kmers_est > ctx->os_kmers_budget || 2 * ovf_seen > ctx->os_ovf_cap) TRY(bins_to_table(ctx));
        else {
            ctx->os_kmers_staged += kmers_est; ctx->kmers_ub_total += kmers_ub;
            TRY(launch_mark(ctx, s, d_offsets, n_reads, n_bases, ctx->cfg.min_seq_len, 1, ctx->d_ctr));
            if (n_bases >= (uint64_t)k) {
                ProfScope ps(ctx, P_EXTRACT_PARTITION, ctx->compute);
                extract_skm_kernel<3, TabAoS><<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
                    d_bases, n_bases, s.d_flags, k, bin_stage(ctx), TabAoS{nullptr, 0}, ctx->d_ctr, nullptr);
                CU_TRY(cudaGetLastError());
            }
            *s.h_snap = 0;
            CU_TRY(cudaMemcpyAsync(s.h_snap, &ctx->d_binctl->ovf_cursor, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->compute));
            CU_TRY(cudaEventRecord(s.ev_done, ctx->compute));
            s.pending = true; s.kmers_submitted_at_end = ctx->kmers_ub_total;
            ctx->dirty = true; ctx->hist_valid = false; ctx->em_valid = false;
            ctx->host_fed = host_fed;
            return MFKC_OK;
        }
    }
    if (ctx->cfg.variant == MFKC_VARIANT_SORT) TRY(sort_variant_reserve(ctx, kmers_ub));
    else { TRY(reserve_slots(ctx, kmers_ub)); table_touched(ctx); }
    if (ctx->cfg.variant == MFKC_VARIANT_HASH) {
        if (ctx->soa && ctx->kmers_since_drain + kmers_ub > 4000000000ull) TRY(drain_regions(ctx));    // u32 counts cannot wrap
        // staging is sized by an ESTIMATE (a full segment only costs speed): bases - reads*(k-1) is exact when
        // every read is at least k-1 long, and 25 % tighter than the guaranteed bound on 150 bp reads
        uint64_t kmers_est = n_bases >= (uint64_t)n_reads * (uint64_t)(k - 1) ? n_bases - (uint64_t)n_reads * (uint64_t)(k - 1) : 0;
        kmers_est = std::max<uint64_t>(kmers_est, kmers_ub / 4);
        TRY(reserve_staging(ctx, kmers_est)); ctx->staged_ub += stage_units(ctx, kmers_est); ctx->kmers_since_drain += kmers_ub;
    }
    ctx->kmers_ub_total += kmers_ub;
    TRY(launch_mark(ctx, s, d_offsets, n_reads, n_bases, ctx->cfg.min_seq_len, 1, ctx->d_ctr));
    if (n_bases >= (uint64_t)k) {
        if (ctx->cfg.variant == MFKC_VARIANT_HASH) {
            ProfScope ps(ctx, P_EXTRACT_PARTITION, ctx->compute);
            static const int stage_mode = getenv("MFKC_STAGE") ? atoi(getenv("MFKC_STAGE")) : 2;
            if (ctx->k128) {                  // 128-bit keys: 32-byte super-k-mer records
                extract_skm128_kernel<0><<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
                    d_bases, n_bases, s.d_flags, k, skm_stage128(ctx), ctx->tab128, ctx->cap, ctx->d_ctr, nullptr);
            } else if (ctx->place) {          // super-k-mer records, minimizer placement
                if (ctx->soa) extract_skm_kernel<0, TabSoAOps><<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
                    d_bases, n_bases, s.d_flags, k, skm_stage(ctx), tab_soa_ops(ctx), ctx->d_ctr, nullptr);
                else extract_skm_kernel<0, TabAoS><<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
                    d_bases, n_bases, s.d_flags, k, skm_stage(ctx), tab_aos(ctx), ctx->d_ctr, nullptr);
            } else if (stage_mode == 0) {     // single keys, shared-memory histogram flavour
                const uint64_t tiles = ((n_bases + 15) / 16 + PT_THREADS - 1) / PT_THREADS;
                const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * 2);
                extract_partition_kernel<<<grid, PT_THREADS, 0, ctx->compute>>>(
                    d_bases, n_bases, s.d_flags, k, region_stage(ctx), ctx->tab, ctx->cap, ctx->d_ctr);
            } else {                          // per-key cursor atomics in L2
                extract_stage_kernel<<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
                    d_bases, n_bases, s.d_flags, k, region_stage(ctx), ctx->tab, ctx->cap, ctx->d_ctr);
            }
        } else if (ctx->cfg.variant == MFKC_VARIANT_HASH_DIRECT) {
            ProfScope ps(ctx, P_EXTRACT_COUNT, ctx->compute);
            SinkTable sink{ctx->tab, ctx->cap, ctx->d_ctr};
            extract_kernel<SinkTable><<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
                d_bases, n_bases, s.d_flags, k, sink, ctx->d_ctr);
        } else {
            ProfScope ps(ctx, P_EXTRACT_KEYS, ctx->compute);
            BucketSink sink{ctx->sv_keys, ctx->d_bucket_base, &ctx->d_ctr->appended, ctx->sv_cap, 1, 0};
            extract_bucket_kernel<0><<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
                d_bases, n_bases, s.d_flags, k, sink, ctx->d_ctr);
        }
        CU_TRY(cudaGetLastError());
    }
    if (ctx->cfg.variant == MFKC_VARIANT_HASH_DIRECT)
        CU_TRY(cudaMemcpyAsync(s.h_snap, &ctx->d_ctr->distinct, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->compute));
    else if (ctx->cfg.variant == MFKC_VARIANT_HASH)
        CU_TRY(cudaMemcpyAsync(s.h_snap, &ctx->d_ctr->kmers, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->compute));
    CU_TRY(cudaEventRecord(s.ev_done, ctx->compute));
    s.pending = ctx->cfg.variant != MFKC_VARIANT_SORT;
    s.kmers_submitted_at_end = ctx->kmers_ub_total;
    ctx->dirty = true; ctx->hist_valid = false; ctx->em_valid = false;
    // Host-fed batches arrive at PCIe speed (~3 ms per million reads), several times slower than the kernels
    // consume them: drain early and often, asynchronously, in the shadow of the copies.  Every drain leaves an
    // exact snapshot of the distinct count (poll_snapshots), which keeps the table bound of reserve_slots tight
    // without ever blocking the submitting thread, and leaves only the last interval for mfkc_flush.
    ctx->host_fed = host_fed;
    if (host_fed && ctx->cfg.variant == MFKC_VARIANT_HASH && ctx->place) {
        static const double frac = getenv("MFKC_DRAIN_EVERY") ? atof(getenv("MFKC_DRAIN_EVERY")) : 0.2;
        const uint64_t expect = ctx->cfg.expected_kmers ? ctx->cfg.expected_kmers : ctx->prev_sample_kmers;
        const uint64_t every = std::max<uint64_t>((uint64_t)(frac * (double)expect), 64ull << 20);
        if (frac > 0 && ctx->kmers_since_drain >= every) TRY(drain_regions(ctx));
    }
    return MFKC_OK;
}

static int p2p_extract_batch(mfkc_ctx *ctx, Staging &s, const uint8_t *d_bases, const uint64_t *d_offsets, uint32_t n_reads, uint64_t n_bases);
static int submit_host(mfkc_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads, bool p2p);

extern "C" int mfkc_submit_reads(mfkc_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads) {
    return submit_host(ctx, bases, offsets, n_reads, false);
}

static int submit_host(mfkc_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads, bool p2p) {
    if (!ctx || (!bases && n_reads) || !offsets) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (n_reads == 0) return MFKC_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    const uint64_t base0 = offsets[0];
    if (offsets[n_reads] < base0) return fail(ctx, MFKC_E_BADARG, "offsets must be non-decreasing");
    const uint64_t n_bases = offsets[n_reads] - base0;
    Staging &s = ctx->st[ctx->next_buf];
    ctx->next_buf = (ctx->next_buf + 1) % N_STAGE;
    TRY(ensure_staging(ctx, s, n_bases, (uint64_t)n_reads + 1, true));
    // the staging buffer is free once the kernels of the batch that used it last are done
    CU_TRY(cudaStreamWaitEvent(s.stream, s.ev_done, 0));
    CU_TRY(cudaMemcpyAsync(s.d_bases, bases + base0, n_bases, cudaMemcpyHostToDevice, s.stream));
    CU_TRY(cudaMemcpyAsync(s.d_offsets, offsets, ((size_t)n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream));
    CU_TRY(cudaEventRecord(s.ev_copy, s.stream));
    CU_TRY(cudaStreamWaitEvent(ctx->compute, s.ev_copy, 0));
    int r = p2p ? p2p_extract_batch(ctx, s, s.d_bases, s.d_offsets, n_reads, n_bases)
                : count_batch_device(ctx, s, s.d_bases, s.d_offsets, n_reads, n_bases, true);
    // the caller may refill its buffers once the copies are done
    CU_TRY(cudaEventSynchronize(s.ev_copy));
    return r;
}

extern "C" int mfkc_submit_reads_device(mfkc_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offsets,
                                        uint32_t n_reads, uint64_t n_bases) {
    if (!ctx || !d_bases || !d_offsets) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (n_reads == 0) return MFKC_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    Staging &s = ctx->st[ctx->next_buf];
    ctx->next_buf = (ctx->next_buf + 1) % N_STAGE;
    TRY(ensure_staging(ctx, s, n_bases, 0, false));
    return count_batch_device(ctx, s, d_bases, d_offsets, n_reads, n_bases);
}

extern "C" int mfkc_flush(mfkc_ctx *ctx) {
    if (!ctx) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(drain_regions(ctx));
    CU_TRY(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->compute));     // rides on the synchronisation below
    TRY(sync_all(ctx));
    if (ctx->cfg.variant != MFKC_VARIANT_SORT) {
        ctx->distinct_ub = ctx->distinct_base = ctx->h_ctr->distinct;
        ctx->kmers_base = ctx->h_ctr->kmers; ctx->recv_since_base = 0;
    }
    ctx->dirty = false;
    if (ctx->h_ctr->bad_chars) return fail(ctx, MFKC_E_FORMAT, "Incorrect nucleotide char in submitted reads (only AaCcGgTt are accepted)");
    if (ctx->h_ctr->overflow) return fail(ctx, MFKC_E_STATE, "internal key buffer overflow");
    return MFKC_OK;
}

// ------------------------------------------------------------------------------------------
// radix sort driver (keys with optional payload), ping-pong between (a) and (b); result in *a
// ------------------------------------------------------------------------------------------
template <typename V, bool HAS_V>
static int radix_sort(mfkc_ctx *ctx, cudaStream_t st, unsigned long long *&a, unsigned long long *&b, V *&va, V *&vb,
                      uint64_t n, int key_bits) {
    if (n < 2) return MFKC_OK;
    const RadixPlan plan = radix_plan(n);
    uint32_t *hist = nullptr; unsigned long long *offs = nullptr, *chunk = nullptr; uint32_t *d_triv = nullptr;
    TMP_ALLOC(hist, plan.hist_bytes);
    TMP_ALLOC(offs, plan.offs_bytes);
    TMP_ALLOC(chunk, plan.chunk_bytes);
    TMP_ALLOC(d_triv, sizeof(uint32_t));
    TRY(stream_after(ctx, st, ctx->compute));
    const size_t smem = rs_scatter_smem_bytes<V, HAS_V>();
    static const int minb = getenv("MFKC_RS_MINB") ? atoi(getenv("MFKC_RS_MINB")) : 4;      // measured on cfg2: 9.8 ms (4) vs 11.3 ms (2) for 120 M records
    CU_TRY(cudaFuncSetAttribute(rs_scatter_kernel<V, HAS_V, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU_TRY(cudaFuncSetAttribute(rs_scatter_kernel<V, HAS_V, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<uint64_t>(plan.n_parts, (uint64_t)ctx->sm_count * 8);
    int rc = MFKC_OK;
    ProfScope ps(ctx, P_SORT, st);
    ctx->prof_launches[P_SORT]--;            // counted per kernel below
    // No host round trip inside the loop: every pass is queued back to back (a pass whose digit is the same for all keys
    // still runs -- it copies in order -- instead of being detected with a read-back and skipped; the number of passes is
    // bounded by key_bits anyway).  A synchronisation per pass cost little on a quiet host and ~1 ms each on a busy one.
    for (int shift = 0; shift < key_bits; shift += 8) {
        rs_hist_kernel<<<grid, RS_THREADS, 0, st>>>(a, n, shift, plan.n_parts, hist);
        rs_chunk_kernel<<<plan.n_chunks, RS_RADIX, 0, st>>>(hist, plan.n_parts, chunk);
        rs_base_kernel<<<1, RS_RADIX, 0, st>>>(chunk, plan.n_chunks, n, d_triv);
        rs_offsets_kernel<<<plan.n_chunks, RS_RADIX, 0, st>>>(hist, plan.n_parts, chunk, offs);
        if (minb == 4) rs_scatter_kernel<V, HAS_V, 4><<<grid, RS_THREADS, smem, st>>>(a, va, n, shift, plan.n_parts, offs, b, vb);
        else rs_scatter_kernel<V, HAS_V, 2><<<grid, RS_THREADS, smem, st>>>(a, va, n, shift, plan.n_parts, offs, b, vb);
        ctx->prof_launches[P_SORT] += 5;
        std::swap(a, b);
        std::swap(va, vb);
    }
    cudaError_t e = cudaGetLastError();
    // the workspace goes back to the pool in stream order: frees are queued on the compute stream behind the sort
    if (e == cudaSuccess && stream_after(ctx, ctx->compute, st) != MFKC_OK) e = cudaErrorUnknown;
    TMP_FREE(hist); TMP_FREE(offs); TMP_FREE(chunk); TMP_FREE(d_triv);
    if (e != cudaSuccess) { ctx->err = std::string("radix sort: ") + cudaGetErrorString(e); return MFKC_E_CUDA; }
    return rc;
}

// ------------------------------------------------------------------------------------------
// sort variant: keys -> sorted run-length (key,count) state
// ------------------------------------------------------------------------------------------
// RLE of a sorted key array (optionally weighted): returns new arrays
static int rle_sorted(mfkc_ctx *ctx, cudaStream_t st, const unsigned long long *keys, const uint32_t *weights, uint64_t n,
                      unsigned long long **out_keys, uint32_t **out_counts, uint64_t *out_n) {
    *out_keys = nullptr; *out_counts = nullptr; *out_n = 0;
    if (n == 0) return MFKC_OK;
    const int grid = grid_for(ctx, n, 256, 8);
    unsigned long long *d_blk = nullptr;
    TMP_ALLOC(d_blk, (size_t)grid * sizeof(unsigned long long));
    ProfScope ps(ctx, P_RLE, st);
    rle_mark_kernel<<<grid, 256, 0, st>>>(keys, n, d_blk);
    std::vector<unsigned long long> h(grid);
    CU_TRY(cudaMemcpyAsync(h.data(), d_blk, (size_t)grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    unsigned long long run = 0;
    for (int i = 0; i < grid; i++) { const unsigned long long c = h[i]; h[i] = run; run += c; }
    CU_TRY(cudaMemcpyAsync(d_blk, h.data(), (size_t)grid * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    unsigned long long *ok = nullptr; uint32_t *oc = nullptr;
    TMP_ALLOC(ok, (size_t)run * sizeof(unsigned long long));
    TMP_ALLOC(oc, (size_t)run * sizeof(uint32_t));
    rle_write_kernel<<<grid, 256, 0, st>>>(keys, weights, n, d_blk, ok, oc);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(st));
    TMP_FREE(d_blk);
    *out_keys = ok; *out_counts = oc; *out_n = run;
    return MFKC_OK;
}

// fold the raw key buffer into the sorted (key,count) state
static int sort_variant_compact(mfkc_ctx *ctx) {
    TRY(sync_all(ctx));
    TRY(read_counters(ctx));
    uint64_t n = ctx->h_ctr->appended;
    if (n > ctx->sv_cap) n = ctx->sv_cap;
    cudaStream_t st = ctx->compute;
    const int key_bits = 2 * ctx->cfg.k;
    if (n) {
        unsigned long long *alt = nullptr; uint32_t *dummy_a = nullptr, *dummy_b = nullptr;
        TMP_ALLOC(alt, (size_t)n * sizeof(unsigned long long));
        unsigned long long *a = ctx->sv_keys, *b = alt;
        int r = radix_sort<uint32_t, false>(ctx, st, a, b, dummy_a, dummy_b, n, key_bits);
        unsigned long long *rk = nullptr; uint32_t *rc = nullptr; uint64_t rn = 0;
        if (r == MFKC_OK) r = rle_sorted(ctx, st, a, nullptr, n, &rk, &rc, &rn);
        TMP_FREE(alt);      // ctx->sv_keys itself (the larger allocation) is kept for the next round
        if (r != MFKC_OK) return r;
        if (ctx->svs_n == 0) {
            TMP_FREE(ctx->svs_keys); TMP_FREE(ctx->svs_counts);
            ctx->svs_keys = rk; ctx->svs_counts = rc; ctx->svs_n = rn;
        } else {
            // merge: concatenate, sort pairs by key, weighted RLE
            const uint64_t m = ctx->svs_n + rn;
            unsigned long long *ck = nullptr, *ck2 = nullptr; uint32_t *cc = nullptr, *cc2 = nullptr;
            TMP_ALLOC(ck, (size_t)m * 8); TMP_ALLOC(ck2, (size_t)m * 8);
            TMP_ALLOC(cc, (size_t)m * 4); TMP_ALLOC(cc2, (size_t)m * 4);
            CU_TRY(cudaMemcpyAsync(ck, ctx->svs_keys, ctx->svs_n * 8, cudaMemcpyDeviceToDevice, st));
            CU_TRY(cudaMemcpyAsync(ck + ctx->svs_n, rk, rn * 8, cudaMemcpyDeviceToDevice, st));
            CU_TRY(cudaMemcpyAsync(cc, ctx->svs_counts, ctx->svs_n * 4, cudaMemcpyDeviceToDevice, st));
            CU_TRY(cudaMemcpyAsync(cc + ctx->svs_n, rc, rn * 4, cudaMemcpyDeviceToDevice, st));
            CU_TRY(cudaStreamSynchronize(st));
            TMP_FREE(rk); TMP_FREE(rc); TMP_FREE(ctx->svs_keys); TMP_FREE(ctx->svs_counts);
            ctx->svs_keys = nullptr; ctx->svs_counts = nullptr; ctx->svs_n = 0;
            r = radix_sort<uint32_t, true>(ctx, st, ck, ck2, cc, cc2, m, key_bits);
            if (r == MFKC_OK) r = rle_sorted(ctx, st, ck, cc, m, &ctx->svs_keys, &ctx->svs_counts, &ctx->svs_n);
            TMP_FREE(ck); TMP_FREE(ck2); TMP_FREE(cc); TMP_FREE(cc2);
            if (r != MFKC_OK) return r;
        }
    }
    CU_TRY(cudaMemsetAsync(&ctx->d_ctr->appended, 0, sizeof(unsigned long long), st));
    CU_TRY(cudaStreamSynchronize(st));
    ctx->sv_ub = 0;
    return MFKC_OK;
}

static int sort_variant_reserve(mfkc_ctx *ctx, uint64_t add) {
    if (ctx->sv_ub + add <= ctx->sv_cap) { ctx->sv_ub += add; return MFKC_OK; }
    // tighten the bound with the real cursor, then grow or fold
    TRY(sync_all(ctx));
    TRY(read_counters(ctx));
    ctx->sv_ub = ctx->h_ctr->appended;
    if (ctx->sv_ub + add > ctx->sv_cap) {
        size_t free_b = 0, total_b = 0;
        CU_TRY(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t want = std::max<uint64_t>((ctx->sv_ub + add) * 2, 1ull << 20);
        // sorting needs a second buffer of the same size plus ~10 % workspace
        const bool can_grow = (double)want * 8.0 * 2.3 + (double)ctx->sv_cap * 8.0 < (double)free_b + (double)ctx->sv_cap * 8.0 &&
                              (double)want * 8.0 * 2.3 < (double)total_b * 0.8;
        if (can_grow) {
            unsigned long long *nk = nullptr;
            CU_TRY(cudaMalloc(&nk, (size_t)want * 8));
            if (ctx->sv_ub) CU_TRY(cudaMemcpy(nk, ctx->sv_keys, (size_t)ctx->sv_ub * 8, cudaMemcpyDeviceToDevice));
            cudaFree(ctx->sv_keys);
            ctx->sv_keys = nk; ctx->sv_cap = want;
        } else {
            TRY(sort_variant_compact(ctx));
            if (add > ctx->sv_cap) return fail(ctx, MFKC_E_OOM, "batch larger than the sort variant's key buffer");
        }
    }
    ctx->sv_ub += add;
    return MFKC_OK;
}

// ------------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------------
static int finalize_counts(mfkc_ctx *ctx) {       // everything submitted is reflected in table / sorted state
    if (ctx->dirty) TRY(mfkc_flush(ctx));
    if (ctx->cfg.variant == MFKC_VARIANT_SORT) {
        TRY(read_counters(ctx));
        if (ctx->h_ctr->appended) TRY(sort_variant_compact(ctx));
    }
    return MFKC_OK;
}

extern "C" int mfkc_stats(mfkc_ctx *ctx, uint64_t stats[6]) {
    if (!ctx || !stats) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(finalize_counts(ctx));
    if (ctx->mode == 1) { const int r = bins_count(ctx, kNoOutput); if (r < 0) return r; }
    TRY(sync_all(ctx));
    TRY(read_counters(ctx));
    stats[0] = ctx->cfg.variant != MFKC_VARIANT_SORT ? ctx->h_ctr->distinct + ctx->h_ctr->bc_distinct : ctx->svs_n;
    stats[1] = ctx->h_ctr->kmers;
    stats[2] = ctx->h_ctr->total_seq; stats[3] = ctx->h_ctr->good_seq;
    stats[4] = ctx->h_ctr->total_len; stats[5] = ctx->h_ctr->good_len;
    return MFKC_OK;
}

// diagnostics of the bin-local mode (see include/mfkc.h)
extern "C" int mfkc_bin_stats(mfkc_ctx *ctx, uint64_t out[8]) {
    if (!ctx || !out) return MFKC_E_BADARG;
    out[0] = (uint64_t)ctx->mode; out[1] = ctx->os_n_bins; out[2] = ctx->os_seg_cap;
    out[3] = ctx->os_heavy_bins; out[4] = ctx->os_heavy_recs; out[5] = ctx->os_splits; out[6] = ctx->os_ovf; out[7] = ctx->os_total_recs;
    return MFKC_OK;
}

static int compute_hist(mfkc_ctx *ctx) {
    if (ctx->hist_valid) return MFKC_OK;
    TRY(finalize_counts(ctx));
    if (ctx->mode == 1) { const int r = bins_count(ctx, kNoOutput); if (r < 0) return r; if (r == 0) return MFKC_OK; }
    cudaStream_t st = ctx->compute;
    CU_TRY(cudaMemsetAsync(ctx->d_hist, 0, MFKC_HIST_BINS * sizeof(unsigned long long), st));
    {
        ProfScope ps(ctx, P_HIST, st);
        if (ctx->k128)
            table128_scan_kernel<<<grid_for(ctx, ctx->cap, 256, 8), 256, 0, st>>>(ctx->tab128, ctx->cap, 0xFFFFFFFFu, ctx->d_hist, nullptr, nullptr, 0, ctx->d_ctr);
        else if (ctx->cfg.variant != MFKC_VARIANT_SORT)
            if (ctx->soa) table_hist_kernel<true><<<grid_for(ctx, ctx->cap, 256, 8), 256, 0, st>>>(nullptr, tab_soa(ctx), ctx->cap, ctx->d_hist);
            else table_hist_kernel<false><<<grid_for(ctx, ctx->cap, 256, 8), 256, 0, st>>>(ctx->tab, TabSoA{nullptr, nullptr, 0}, ctx->cap, ctx->d_hist);
        else if (ctx->svs_n)
            pairs_hist_kernel<<<grid_for(ctx, ctx->svs_n, 256, 8), 256, 0, st>>>(ctx->svs_counts, ctx->svs_n, ctx->d_hist);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(ctx->h_hist, ctx->d_hist, MFKC_HIST_BINS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    ctx->hist_valid = true;
    return MFKC_OK;
}

extern "C" int mfkc_histogram(mfkc_ctx *ctx, uint64_t hist[MFKC_HIST_BINS]) {
    if (!ctx || !hist) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(compute_hist(ctx));
    memcpy(hist, ctx->h_hist, MFKC_HIST_BINS * sizeof(uint64_t));
    return MFKC_OK;
}

// order-preserving filter of the sorted state (sort variant): counts > threshold
__global__ void __launch_bounds__(256)
select_mark_kernel(const uint32_t *__restrict__ counts, uint64_t n, uint32_t threshold, unsigned long long *__restrict__ blk) {
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint64_t per_block = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = per_block * blockIdx.x;
    const uint64_t hi = lo + per_block < n ? lo + per_block : n;
    uint32_t local = 0;
    for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) local += counts[i] > threshold ? 1u : 0u;
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) blk[blockIdx.x] = s_cnt;
}
__global__ void __launch_bounds__(256)
select_write_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ counts, uint64_t n,
                    uint32_t threshold, const unsigned long long *__restrict__ blk_base,
                    unsigned long long *__restrict__ out_keys, uint16_t *__restrict__ out_counts) {
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_base;
    const uint64_t per_block = (n + gridDim.x - 1) / gridDim.x;
    const uint64_t lo = per_block * blockIdx.x;
    const uint64_t hi = lo + per_block < n ? lo + per_block : n;
    if (threadIdx.x == 0) s_base = blk_base[blockIdx.x];
    __syncthreads();
    for (uint64_t start = lo; start < hi; start += blockDim.x) {
        const uint64_t i = start + threadIdx.x;
        const uint32_t c = i < hi ? counts[i] : 0u;
        const bool good = i < hi && c > threshold;
        const uint32_t m = __ballot_sync(0xffffffffu, good);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int wv = 0; wv < 8; wv++) { const uint32_t cc = s_warp[wv]; if (wv < (int)(threadIdx.x >> 5)) before += cc; total += cc; }
        if (good) {
            const uint64_t at = s_base + before + __popc(m & mfkc::lanemask_lt());
            out_keys[at] = keys[i];
            out_counts[at] = (uint16_t)(c < MAX_COUNT ? c : MAX_COUNT);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
}

extern "C" int mfkc_emit_begin(mfkc_ctx *ctx, int32_t threshold, uint64_t *n_good) {
    if (!ctx) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(finalize_counts(ctx));
    free_emit(ctx);
    // entries with value > threshold (src/io/IOUtils.java:61); stored counts are 1..32767, so every
    // negative threshold behaves like 0
    const uint32_t thr_u = threshold < 0 ? 0u : (uint32_t)threshold;
    cudaStream_t st = ctx->compute;
    uint64_t good = 0;
    if (ctx->k128) {
        TRY(read_counters(ctx));
        const uint64_t distinct = ctx->h_ctr->distinct;
        CU_TRY(cudaMemsetAsync(ctx->d_hist, 0, MFKC_HIST_BINS * sizeof(unsigned long long), st));
        CU_TRY(cudaMemsetAsync(&ctx->d_ctr->n_good, 0, sizeof(unsigned long long), st));
        unsigned long long *k1 = nullptr, *k2 = nullptr; HiCount *v1 = nullptr, *v2 = nullptr;
        TMP_ALLOC(k1, (size_t)std::max<uint64_t>(distinct, 1) * 8);
        TMP_ALLOC(v1, (size_t)std::max<uint64_t>(distinct, 1) * sizeof(HiCount));
        {
            ProfScope ps(ctx, P_COMPACT, st);
            table128_scan_kernel<<<grid_for(ctx, ctx->cap, 256, 8), 256, 0, st>>>(ctx->tab128, ctx->cap, thr_u, ctx->d_hist, k1, v1, distinct, ctx->d_ctr);
        }
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(ctx->h_hist, ctx->d_hist, MFKC_HIST_BINS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(&ctx->h_ctr->n_good, &ctx->d_ctr->n_good, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        ctx->hist_valid = true;
        good = ctx->h_ctr->n_good;
        if (good) {
            TMP_ALLOC(k2, (size_t)good * 8);
            TMP_ALLOC(v2, (size_t)good * sizeof(HiCount));
            int r = radix_sort<HiCount, true>(ctx, st, k1, k2, v1, v2, good, 64);                       // by the low word ...
            if (r == MFKC_OK) {
                swap_key_words_kernel<<<grid_for(ctx, good, 256, 8), 256, 0, st>>>(k1, v1, good);
                r = radix_sort<HiCount, true>(ctx, st, k1, k2, v1, v2, good, 2 * ctx->cfg.k - 64);      // ... then (stable) by the high word
            }
            if (r != MFKC_OK) { TMP_FREE(k1); TMP_FREE(k2); TMP_FREE(v1); TMP_FREE(v2); return r; }
            TMP_ALLOC(ctx->em_records, (size_t)good * 18);
            {
                ProfScope ps(ctx, P_RECORDS, st);
                records128_kernel<<<grid_for(ctx, good, 256, 8), 256, 0, st>>>(k1, v1, good, reinterpret_cast<uint16_t *>(ctx->em_records));
            }
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaStreamSynchronize(st));
        }
        TMP_FREE(k1); TMP_FREE(k2); TMP_FREE(v1); TMP_FREE(v2);
        ctx->em_n = good; ctx->em_cursor = 0; ctx->em_valid = true;
        if (n_good) *n_good = good;
        return MFKC_OK;
    }
    if (ctx->cfg.variant != MFKC_VARIANT_SORT) {
        unsigned long long *k1 = nullptr; uint16_t *c1 = nullptr;
        bool counted = false;
        if (ctx->mode == 1) {
            // bin-local mode: the count pass itself selects the entries with count > threshold (unsorted)
            const int r = bins_count(ctx, thr_u);
            if (r < 0) return r;
            if (r == 0) {
                counted = true;
                static const bool low = getenv("MFKC_EMIT_PRIORITY") == nullptr || atoi(getenv("MFKC_EMIT_PRIORITY")) != 0;
                if (low) st = ctx->emit;
                good = ctx->os_n_good;
                k1 = ctx->os_keys; c1 = ctx->os_counts; ctx->os_keys = nullptr; ctx->os_counts = nullptr;
                if (!good) { TMP_FREE(k1); TMP_FREE(c1); k1 = nullptr; c1 = nullptr; }
            }
        }
        if (!counted) {
            // one pass over the table: histogram + compaction (buffers sized by the exact distinct count)
            TRY(read_counters(ctx));
            const uint64_t distinct = ctx->h_ctr->distinct;
            if (distinct) {
                TMP_ALLOC(k1, (size_t)distinct * 8);
                if (cudaMallocAsync((void **)&c1, (size_t)distinct * 2, st) != cudaSuccess) { cudaGetLastError(); TMP_FREE(k1); return fail(ctx, MFKC_E_OOM, "cannot allocate the emit buffers"); }
                cudaError_t e = cudaMemsetAsync(ctx->d_hist, 0, MFKC_HIST_BINS * sizeof(unsigned long long), st);
                if (e == cudaSuccess) e = cudaMemsetAsync(&ctx->d_ctr->n_good, 0, sizeof(unsigned long long), st);
                if (e == cudaSuccess) {
                    ProfScope ps(ctx, P_COMPACT, st);
                    if (ctx->soa) table_scan_kernel<true><<<grid_for(ctx, (ctx->cap + 7) / 8, 256, 8), 256, 0, st>>>(nullptr, tab_soa(ctx), ctx->cap, thr_u, ctx->d_hist, k1, c1, distinct, ctx->d_ctr);
                    else table_scan_kernel<false><<<grid_for(ctx, (ctx->cap + 7) / 8, 256, 8), 256, 0, st>>>(ctx->tab, TabSoA{nullptr, nullptr, 0}, ctx->cap, thr_u, ctx->d_hist, k1, c1, distinct, ctx->d_ctr);
                    e = cudaGetLastError();
                }
                if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_hist, ctx->d_hist, MFKC_HIST_BINS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaMemcpyAsync(&ctx->h_ctr->n_good, &ctx->d_ctr->n_good, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) { TMP_FREE(k1); TMP_FREE(c1); ctx->err = std::string("emit: ") + cudaGetErrorString(e); return MFKC_E_CUDA; }
                ctx->hist_valid = true;
                good = ctx->h_ctr->n_good;
                if (!good) { TMP_FREE(k1); TMP_FREE(c1); k1 = nullptr; c1 = nullptr; }
            } else {
                TRY(compute_hist(ctx));
            }
        }
        if (good) {
            unsigned long long *k2 = nullptr; uint16_t *c2 = nullptr;
            cudaError_t e = cudaMallocAsync((void **)&k2, (size_t)good * 8, st);
            if (e == cudaSuccess) e = cudaMallocAsync((void **)&c2, (size_t)good * 2, st);
            if (e != cudaSuccess) { cudaGetLastError(); TMP_FREE(k1); TMP_FREE(c1); TMP_FREE(k2); return fail(ctx, MFKC_E_OOM, "cannot allocate the sort buffers"); }
            const int r = radix_sort<uint16_t, true>(ctx, st, k1, k2, c1, c2, good, 2 * ctx->cfg.k);
            if (r != MFKC_OK) { TMP_FREE(k1); TMP_FREE(k2); TMP_FREE(c1); TMP_FREE(c2); return r; }
            ctx->em_keys = k1; ctx->em_counts = c1;       // radix_sort leaves the result in (k1, c1)
            TMP_FREE(k2); TMP_FREE(c2);
        }
    } else {
        TRY(compute_hist(ctx));
        for (int c = 1; c < MFKC_HIST_BINS; c++) if ((uint32_t)c > thr_u) good += ctx->h_hist[c];
        if (good) {
            TMP_ALLOC(ctx->em_keys, (size_t)good * 8);
            TMP_ALLOC(ctx->em_counts, (size_t)good * 2);
            const int grid = grid_for(ctx, ctx->svs_n, 256, 8);
            unsigned long long *d_blk = nullptr;
            TMP_ALLOC(d_blk, (size_t)grid * 8);
            ProfScope ps(ctx, P_COMPACT, st);
            select_mark_kernel<<<grid, 256, 0, st>>>(ctx->svs_counts, ctx->svs_n, thr_u, d_blk);
            std::vector<unsigned long long> h(grid);
            CU_TRY(cudaMemcpyAsync(h.data(), d_blk, (size_t)grid * 8, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            unsigned long long run = 0;
            for (int i = 0; i < grid; i++) { const unsigned long long c = h[i]; h[i] = run; run += c; }
            CU_TRY(cudaMemcpyAsync(d_blk, h.data(), (size_t)grid * 8, cudaMemcpyHostToDevice, st));
            select_write_kernel<<<grid, 256, 0, st>>>(ctx->svs_keys, ctx->svs_counts, ctx->svs_n, thr_u, d_blk,
                                                     ctx->em_keys, ctx->em_counts);
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaStreamSynchronize(st));
            TMP_FREE(d_blk);
            if (run != good) return fail(ctx, MFKC_E_STATE, "internal: selection count mismatch");
        }
    }
    ctx->em_n = good;
    if (good) {
        TMP_ALLOC(ctx->em_records, (size_t)good * 10);
        TRY(stream_after(ctx, st, ctx->compute));
        {
            ProfScope ps(ctx, P_RECORDS, st);
            records_kernel<<<grid_for(ctx, good, 256, 8), 256, 0, st>>>(ctx->em_keys, ctx->em_counts, good,
                                                                      reinterpret_cast<uint16_t *>(ctx->em_records));
        }
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaStreamSynchronize(st));
    }
    ctx->em_cursor = 0; ctx->em_valid = true;
    if (n_good) *n_good = good;
    return MFKC_OK;
}

extern "C" int mfkc_emit_next(mfkc_ctx *ctx, uint8_t *out, size_t cap, size_t *written) {
    if (!ctx || !written) return MFKC_E_BADARG;
    if (!ctx->em_valid) return fail(ctx, MFKC_E_STATE, "mfkc_emit_next without mfkc_emit_begin");
    CU_TRY(cudaSetDevice(ctx->device));
    const uint64_t left = ctx->em_n - ctx->em_cursor;
    const size_t rs = ctx->k128 ? 18 : 10;
    uint64_t take = std::min<uint64_t>(left, cap / rs);
    if (take && !out) return MFKC_E_BADARG;
    if (take) {
        // on a stream of its own (not the legacy default stream): the copy back overlaps another context's host-to-device copies
        CU_TRY(cudaMemcpyAsync(out, ctx->em_records + ctx->em_cursor * rs, (size_t)take * rs, cudaMemcpyDeviceToHost, ctx->emit));
        CU_TRY(cudaStreamSynchronize(ctx->emit));
    }
    ctx->em_cursor += take;
    *written = (size_t)take * rs;
    return MFKC_OK;
}

extern "C" int mfkc_emit_device(mfkc_ctx *ctx, const uint64_t **d_keys, const uint16_t **d_counts, uint64_t *n) {
    if (!ctx) return MFKC_E_BADARG;
    if (!ctx->em_valid) return fail(ctx, MFKC_E_STATE, "mfkc_emit_device without mfkc_emit_begin");
    if (ctx->k128) return fail(ctx, MFKC_E_STATE, "mfkc_emit_device serves 64-bit keys only");
    if (d_keys) *d_keys = reinterpret_cast<const uint64_t *>(ctx->em_keys);
    if (d_counts) *d_counts = ctx->em_counts;
    if (n) *n = ctx->em_n;
    return MFKC_OK;
}

// ------------------------------------------------------------------------------------------
// hash-range sharding helpers
// ------------------------------------------------------------------------------------------
extern "C" uint32_t mfkc_owner_shard(uint64_t key, uint32_t n_shards) { return n_shards > 1 ? owner_shard(key, n_shards) : 0u; }

extern "C" int mfkc_extract_bucketed(mfkc_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offsets, uint32_t n_reads,
                                     uint64_t n_bases, uint64_t *d_keys_out, uint64_t cap_keys, uint64_t *bucket_counts) {
    if (!ctx || !d_bases || !d_offsets || !d_keys_out || !bucket_counts) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (ctx->k128) return fail(ctx, MFKC_E_STATE, "this flavour of the shard exchange serves k <= 31; use mfkc_p2p_* for 128-bit keys");
    const uint32_t ns = ctx->cfg.n_shards > 1 ? (uint32_t)ctx->cfg.n_shards : 1u;
    CU_TRY(cudaSetDevice(ctx->device));
    Staging &s = ctx->st[0];
    TRY(ensure_staging(ctx, s, n_bases, 0, false));
    for (uint32_t i = 0; i < ns; i++) bucket_counts[i] = 0;
    if (n_reads == 0) return MFKC_OK;
    const int k = ctx->cfg.k;
    TRY(launch_mark(ctx, s, d_offsets, n_reads, n_bases, ctx->cfg.min_seq_len, 1, ctx->d_ctr));
    if (n_bases < (uint64_t)k) { CU_TRY(cudaStreamSynchronize(ctx->compute)); return MFKC_OK; }
    const int grid = extract_grid(ctx, n_bases);
    // pass 1: bucket sizes
    CU_TRY(cudaMemsetAsync(ctx->d_bucket_cursor, 0, 64 * sizeof(unsigned long long), ctx->compute));
    {
        ProfScope ps(ctx, P_EXTRACT_KEYS, ctx->compute);
        BucketSink sink{reinterpret_cast<unsigned long long *>(d_keys_out), ctx->d_bucket_base, ctx->d_bucket_cursor, cap_keys, ns, 1};
        extract_bucket_kernel<0><<<grid, EX_THREADS, 0, ctx->compute>>>(d_bases, n_bases, s.d_flags, k, sink, ctx->d_ctr);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(ctx->h_bucket, ctx->d_bucket_cursor, ns * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->compute));
    CU_TRY(cudaStreamSynchronize(ctx->compute));
    uint64_t base[64], run = 0;
    for (uint32_t i = 0; i < ns; i++) { bucket_counts[i] = ctx->h_bucket[i]; base[i] = run; run += ctx->h_bucket[i]; }
    if (run > cap_keys) return fail(ctx, MFKC_E_BADARG, "d_keys_out too small for this batch");
    CU_TRY(cudaMemcpyAsync(ctx->d_bucket_base, base, ns * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->compute));
    CU_TRY(cudaMemsetAsync(ctx->d_bucket_cursor, 0, 64 * sizeof(unsigned long long), ctx->compute));
    // pass 2: write grouped keys
    {
        ProfScope ps(ctx, P_EXTRACT_KEYS, ctx->compute);
        BucketSink sink{reinterpret_cast<unsigned long long *>(d_keys_out), ctx->d_bucket_base, ctx->d_bucket_cursor, cap_keys, ns, 0};
        extract_bucket_kernel<0><<<grid, EX_THREADS, 0, ctx->compute>>>(d_bases, n_bases, s.d_flags, k, sink, ctx->d_ctr);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(ctx->compute));
    // restore the sort-variant base (bucket 0 starts at 0)
    CU_TRY(cudaMemsetAsync(ctx->d_bucket_base, 0, 64 * sizeof(uint64_t), ctx->compute));
    return MFKC_OK;
}

extern "C" int mfkc_count_keys_device(mfkc_ctx *ctx, const uint64_t *d_keys, uint64_t n) {
    if (!ctx || (!d_keys && n)) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (ctx->cfg.variant == MFKC_VARIANT_SORT) return fail(ctx, MFKC_E_STATE, "mfkc_count_keys_device needs a hash variant");
    if (ctx->k128) return fail(ctx, MFKC_E_STATE, "this flavour of the shard exchange serves k <= 31; use mfkc_p2p_* for 128-bit keys");
    if (n == 0) return MFKC_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(ensure_table_mode(ctx));
    TRY(reserve_slots(ctx, n));
    const bool stage_keys = ctx->cfg.variant == MFKC_VARIANT_HASH && !ctx->place;
    if (stage_keys) { TRY(reserve_staging(ctx, n)); ctx->staged_ub += n; }
    ctx->kmers_ub_total += n; ctx->recv_since_base += n;
    Staging &s = ctx->st[ctx->next_buf];
    ctx->next_buf = (ctx->next_buf + 1) % N_STAGE;
    if (stage_keys) {
        ProfScope ps(ctx, P_EXTRACT_PARTITION, ctx->compute);
        const uint64_t tiles = (n + PT_THREADS * 16 - 1) / (PT_THREADS * 16);
        const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * 2);
        partition_keys_kernel<<<grid, PT_THREADS, 0, ctx->compute>>>(
            reinterpret_cast<const unsigned long long *>(d_keys), n, region_stage(ctx), ctx->tab, ctx->cap, ctx->d_ctr);
    } else {
        ProfScope ps(ctx, P_COUNT_KEYS, ctx->compute);
        if (ctx->soa) {
            if (ctx->kmers_since_drain + n > 4000000000ull) TRY(drain_regions(ctx));
            ctx->kmers_since_drain += n;
            count_keys_soa_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->compute>>>(
                reinterpret_cast<const unsigned long long *>(d_keys), n, tab_soa(ctx), table_geom(ctx), ctx->d_ctr);
        } else
            count_keys_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->compute>>>(
                reinterpret_cast<const unsigned long long *>(d_keys), n, ctx->tab, table_geom(ctx), ctx->d_ctr);
    }
    CU_TRY(cudaGetLastError());
    if (ctx->cfg.variant == MFKC_VARIANT_HASH_DIRECT)
        CU_TRY(cudaMemcpyAsync(s.h_snap, &ctx->d_ctr->distinct, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->compute));
    CU_TRY(cudaEventRecord(s.ev_done, ctx->compute));
    s.pending = ctx->cfg.variant == MFKC_VARIANT_HASH_DIRECT; s.kmers_submitted_at_end = ctx->kmers_ub_total;
    ctx->dirty = true; ctx->hist_valid = false; ctx->em_valid = false;
    return MFKC_OK;
}

// ---- super-k-mer flavour of the shard exchange (the default for MFKC_VARIANT_HASH) ----------
// Send side: records bucketed by owner shard.  d_recs_out holds n_shards segments of seg_cap records
// (16 bytes each); rec_counts[s] / kmer_counts[s] (host) = records / k-mer instances for shard s.
// Returns 1 (not an error) when a segment overflowed: nothing may be used, retry with fewer reads.
extern "C" int mfkc_skm_extract_bucketed(mfkc_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offsets, uint32_t n_reads,
                                         uint64_t n_bases, void *d_recs_out, uint64_t seg_cap, uint64_t *rec_counts,
                                         uint64_t *kmer_counts) {
    if (!ctx || !d_bases || !d_offsets || !d_recs_out || !rec_counts || !kmer_counts) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (ctx->k128) return fail(ctx, MFKC_E_STATE, "this flavour of the shard exchange serves k <= 31; use mfkc_p2p_* for 128-bit keys");
    const uint32_t ns = ctx->cfg.n_shards > 1 ? (uint32_t)ctx->cfg.n_shards : 1u;
    if (seg_cap == 0 || seg_cap > 0x7fffffffull) return fail(ctx, MFKC_E_BADARG, "bad segment capacity");
    CU_TRY(cudaSetDevice(ctx->device));
    Staging &s = ctx->st[0];
    TRY(ensure_staging(ctx, s, n_bases, 0, false));
    for (uint32_t i = 0; i < ns; i++) { rec_counts[i] = 0; kmer_counts[i] = 0; }
    if (n_reads == 0) return MFKC_OK;
    const int k = ctx->cfg.k;
    // the read statistics are taken once per read, also when the caller has to retry a batch in halves:
    // statistics are accumulated only by the successful attempt (count_stats below), so snapshot + restore
    CU_TRY(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->compute));
    TRY(launch_mark(ctx, s, d_offsets, n_reads, n_bases, ctx->cfg.min_seq_len, 1, ctx->d_ctr));
    if (n_bases < (uint64_t)k) { CU_TRY(cudaStreamSynchronize(ctx->compute)); return MFKC_OK; }
    unsigned int *d_cur = reinterpret_cast<unsigned int *>(ctx->d_bucket_cursor);        // 64 x u64 scratch: cursors (u32) ...
    unsigned long long *d_kc = reinterpret_cast<unsigned long long *>(ctx->d_bucket_base);                                       // ... and k-mer counts (u64)
    CU_TRY(cudaMemsetAsync(ctx->d_bucket_cursor, 0, 64 * sizeof(unsigned long long), ctx->compute));
    CU_TRY(cudaMemsetAsync(ctx->d_bucket_base, 0, 64 * sizeof(uint64_t), ctx->compute));
    SkmStage st{};
    st.recs = reinterpret_cast<uint4 *>(d_recs_out); st.cursor = d_cur; st.seg_cap = seg_cap; st.n_regions = ns; st.region_shift = 0; st.win = 0;
    {
        ProfScope ps(ctx, P_EXTRACT_BUCKET, ctx->compute);
        if (ns <= 8)
            extract_skm_owner8_kernel<<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
                d_bases, n_bases, s.d_flags, k, st, ctx->d_ctr, d_kc);
        else
            extract_skm_kernel<1, TabAoS><<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
                d_bases, n_bases, s.d_flags, k, st, TabAoS{nullptr, 0}, ctx->d_ctr, d_kc);
    }
    CU_TRY(cudaGetLastError());
    unsigned int h_cur[64];
    CU_TRY(cudaMemcpyAsync(h_cur, d_cur, ns * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->compute));
    CU_TRY(cudaMemcpyAsync(ctx->h_bucket, d_kc, ns * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->compute));
    CU_TRY(cudaStreamSynchronize(ctx->compute));
    bool overflow = false;
    for (uint32_t i = 0; i < ns; i++) { rec_counts[i] = h_cur[i]; kmer_counts[i] = ctx->h_bucket[i]; if (h_cur[i] > seg_cap) overflow = true; }
    CU_TRY(cudaMemsetAsync(ctx->d_bucket_base, 0, 64 * sizeof(uint64_t), ctx->compute)); // bucket 0 of the sort variant starts at 0
    if (overflow) {
        // undo this attempt's statistics; the caller re-submits the same reads in smaller pieces
        CU_TRY(cudaMemcpy(ctx->d_ctr, ctx->h_ctr, sizeof(Counters), cudaMemcpyHostToDevice));
        return 1;
    }
    return MFKC_OK;
}

// Receive side: n records (device pointer) with n_kmers k-mer instances in total are filed into this
// context's region staging; they are counted at the next drain (mfkc_flush at the latest).
extern "C" int mfkc_skm_count_device(mfkc_ctx *ctx, const void *d_recs, uint64_t n_recs, uint64_t n_kmers) {
    if (!ctx || (!d_recs && n_recs)) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (ctx->cfg.variant != MFKC_VARIANT_HASH || !ctx->place || ctx->k128) return fail(ctx, MFKC_E_STATE, "mfkc_skm_count_device needs the region-blocked hash variant with k <= 31");
    if (n_recs == 0) return MFKC_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(ensure_table_mode(ctx));
    TRY(reserve_slots(ctx, n_kmers));
    if (ctx->soa && ctx->kmers_since_drain + n_kmers > 4000000000ull) TRY(drain_regions(ctx));
    ctx->kmers_since_drain += n_kmers;
    // reserve staging by records (2 units each), not by the k-mer estimate
    {
        const uint64_t units = 2 * n_recs;
        if (!ctx->rb_keys || ctx->staged_ub + units > std::min<uint64_t>(ctx->rb_cap, 4000000000ull)) TRY(reserve_staging(ctx, n_kmers));
        ctx->staged_ub += units;
    }
    ctx->kmers_ub_total += n_kmers; ctx->recv_since_base += n_kmers;
    {
        // on the aux stream: the next round's extraction (compute stream) overlaps this kernel
        ProfScope ps(ctx, P_EXTRACT_PARTITION, ctx->aux);
        if (ctx->soa) skm_restage_kernel<TabSoAOps><<<grid_for(ctx, n_recs, 256, 4), 256, 0, ctx->aux>>>(
            reinterpret_cast<const uint4 *>(d_recs), n_recs, ctx->cfg.k, skm_stage(ctx), tab_soa_ops(ctx), ctx->d_ctr);
        else skm_restage_kernel<TabAoS><<<grid_for(ctx, n_recs, 256, 4), 256, 0, ctx->aux>>>(
            reinterpret_cast<const uint4 *>(d_recs), n_recs, ctx->cfg.k, skm_stage(ctx), tab_aos(ctx), ctx->d_ctr);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(ctx->ev_aux, ctx->aux));
    ctx->aux_pending = true;
    ctx->dirty = true; ctx->hist_valid = false; ctx->em_valid = false;
    return MFKC_OK;
}

extern "C" int mfkc_skm_count_wait(mfkc_ctx *ctx) {
    if (!ctx) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaStreamSynchronize(ctx->aux));
    return MFKC_OK;
}

// ---- peer-memory flavour of the shard exchange (see include/mfkc.h and drain_p2p_kernel) ----
static int p2p_check(mfkc_ctx *ctx) {
    if (!ctx) return MFKC_E_BADARG;
    if (ctx->cfg.variant != MFKC_VARIANT_HASH || !ctx->place || ctx->soa)
        return fail(ctx, MFKC_E_STATE, "the peer-memory exchange needs the region-blocked hash variant");
    if (ctx->cfg.n_shards < 1 || ctx->cfg.n_shards > P2P_MAX_PEERS) return fail(ctx, MFKC_E_STATE, "the peer-memory exchange serves 1..16 shards");
    return MFKC_OK;
}

extern "C" int mfkc_p2p_stage_create(mfkc_ctx *ctx, uint32_t log2_buckets, uint64_t seg_cap) {
    TRY(p2p_check(ctx));
    if (log2_buckets > 16 || seg_cap == 0 || seg_cap > 0x7fffffffull) return fail(ctx, MFKC_E_BADARG, "bad p2p staging geometry");
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(sync_all(ctx));
    cudaFree(ctx->p2p_recs); cudaFree(ctx->p2p_cursor); cudaFree(ctx->p2p_kc);
    ctx->p2p_recs = nullptr; ctx->p2p_cursor = nullptr; ctx->p2p_kc = nullptr;
    const uint64_t n_seg = (uint64_t)std::max(1, ctx->cfg.n_shards) << log2_buckets;
    // plain cudaMalloc: CUDA IPC cannot export memory of the stream-ordered pool
    if (big_alloc(ctx, (void **)&ctx->p2p_recs, n_seg * seg_cap * sizeof(uint4) * (ctx->k128 ? 2 : 1)) != cudaSuccess) return fail(ctx, MFKC_E_OOM, "cannot allocate the p2p staging buffer");
    CU_TRY(cudaMalloc(&ctx->p2p_cursor, n_seg * sizeof(unsigned int)));
    CU_TRY(cudaMalloc(&ctx->p2p_kc, P2P_MAX_PEERS * sizeof(unsigned long long)));
    CU_TRY(cudaMemset(ctx->p2p_cursor, 0, n_seg * sizeof(unsigned int)));
    CU_TRY(cudaMemset(ctx->p2p_kc, 0, P2P_MAX_PEERS * sizeof(unsigned long long)));
    CU_TRY(cudaDeviceSynchronize());                                          // peers write these once the handles are out
    ctx->p2p_seg_cap = seg_cap; ctx->p2p_log2 = (int)log2_buckets;
    ctx->p2p_bins = false; ctx->p2p_n_cursor = n_seg;
    return MFKC_OK;
}

// Geometry of the bin-local peer-memory staging, the same on every rank: a shard receives about what a rank sends
// (kmers_per_rank instances); its bins are sized so that the distinct k-mers of one bin load the shared-memory table to
// ~0.45 (instances_per_distinct: 0 = unknown, 3 assumed); every (sender, owner, bin) segment gets `slack` (0 = 2.0)
// times its expected records (one record per 16-base word + one per minimizer change), the overflow list 6 %.
extern "C" int mfkc_p2p_bin_geometry(uint64_t kmers_per_rank, uint32_t n_shards, int k, double instances_per_distinct, double slack,
                                     uint32_t *bins_per_shard, uint64_t *seg_cap, uint64_t *ovf_cap) {
    if (!bins_per_shard || !seg_cap || !ovf_cap || n_shards < 1 || k < 1) return MFKC_E_BADARG;
    const double R = instances_per_distinct >= 1.0 ? instances_per_distinct : 3.0;
    const double sl = slack > 0 ? slack : 2.0;
    const double kmers = (double)kmers_per_rank * 1.03 + 4096.0;
    uint64_t bins = (uint64_t)(kmers / R / (0.45 * (double)(1u << BC_LOG2S))) + 1;
    bins = std::min<uint64_t>(std::max<uint64_t>(bins, 16), 1ull << 24);
    const int w = k - bin_minimizer_len(k, bins * n_shards) + 1;
    const double rpk = std::min(1.0, (1.0 / 16 + 2.0 / (w + 1)) * 1.2);
    const double recs = kmers * rpk;
    *bins_per_shard = (uint32_t)bins;
    *seg_cap = (uint64_t)(recs / ((double)n_shards * (double)bins) * sl) + 64;
    *ovf_cap = std::max<uint64_t>((uint64_t)(recs * 0.06), 1ull << 16);
    return MFKC_OK;
}

// Staging for the bin-local count over peer memory: n_shards x bins_per_shard segments of seg_cap records, then the
// overflow list (ovf_cap records); cursors: one per segment + the overflow cursor.  Every rank passes the same geometry.
extern "C" int mfkc_p2p_stage_create_bins(mfkc_ctx *ctx, uint32_t bins_per_shard, uint64_t seg_cap, uint64_t ovf_cap) {
    TRY(p2p_check(ctx));
    if (!ctx->bins_ok) return fail(ctx, MFKC_E_STATE, "the bin-local exchange needs MFKC_VARIANT_HASH with k <= 31 and no pinned table geometry");
    if (bins_per_shard == 0 || bins_per_shard > (1u << 24) || seg_cap == 0 || seg_cap > 0x7fffffffull || ovf_cap == 0 || ovf_cap > 0x7fffffffull)
        return fail(ctx, MFKC_E_BADARG, "bad p2p staging geometry");
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(sync_all(ctx));
    cudaFree(ctx->p2p_recs); cudaFree(ctx->p2p_cursor); cudaFree(ctx->p2p_kc);
    ctx->p2p_recs = nullptr; ctx->p2p_cursor = nullptr; ctx->p2p_kc = nullptr;
    const uint64_t n_seg = (uint64_t)std::max(1, ctx->cfg.n_shards) * bins_per_shard;
    ovf_cap = (ovf_cap + 31) / 32 * 32;                                       // whole chunks; the tag table follows the pool
    if (big_alloc(ctx, (void **)&ctx->p2p_recs, n_seg * seg_cap * sizeof(uint4) + ovf_bytes(ovf_cap)) != cudaSuccess)
        return fail(ctx, MFKC_E_OOM, "cannot allocate the p2p staging buffer");
    CU_TRY(cudaMalloc(&ctx->p2p_cursor, (n_seg + 1) * sizeof(unsigned int)));
    CU_TRY(cudaMalloc(&ctx->p2p_kc, P2P_MAX_PEERS * sizeof(unsigned long long)));
    CU_TRY(cudaMemset(ctx->p2p_cursor, 0, (n_seg + 1) * sizeof(unsigned int)));
    CU_TRY(cudaMemset(ctx->p2p_kc, 0, P2P_MAX_PEERS * sizeof(unsigned long long)));
    CU_TRY(cudaDeviceSynchronize());                                          // peers write these once the handles are out
    ctx->p2p_seg_cap = seg_cap; ctx->p2p_log2 = 0; ctx->p2p_bins = true; ctx->p2p_B = bins_per_shard; ctx->p2p_ovf_cap = ovf_cap;
    ctx->os_mlen = bin_minimizer_len(ctx->cfg.k, n_seg);         // the same on every rank: a function of the geometry
    ctx->p2p_n_cursor = n_seg + 1;
    ctx->os_n_bins = bins_per_shard; ctx->os_seg_cap = seg_cap; ctx->os_ovf_cap = ovf_cap;
    if (ctx->heavy_cap < (uint64_t)bins_per_shard + 4096) {
        cudaFree(ctx->d_heavy); ctx->d_heavy = nullptr;
        ctx->heavy_cap = bins_per_shard + bins_per_shard / 4 + 4096;
        CU_TRY(cudaMalloc(&ctx->d_heavy, (size_t)ctx->heavy_cap * sizeof(HeavyEnt)));
    }
    return MFKC_OK;
}

extern "C" int mfkc_p2p_export(mfkc_ctx *ctx, uint8_t handles[128]) {
    TRY(p2p_check(ctx));
    if (!handles || !ctx->p2p_recs) return fail(ctx, MFKC_E_STATE, "mfkc_p2p_export before mfkc_p2p_stage_create");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    CU_TRY(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, ctx->p2p_recs)); memcpy(handles, &h, 64);
    CU_TRY(cudaIpcGetMemHandle(&h, ctx->p2p_cursor)); memcpy(handles + 64, &h, 64);
    return MFKC_OK;
}

extern "C" int mfkc_p2p_attach(mfkc_ctx *ctx, uint32_t rank, const uint8_t handles[128]) {
    TRY(p2p_check(ctx));
    if (rank >= (uint32_t)std::max(1, ctx->cfg.n_shards)) return fail(ctx, MFKC_E_BADARG, "bad peer rank");
    CU_TRY(cudaSetDevice(ctx->device));
    if ((int)rank == ctx->cfg.shard_id) {
        if (!ctx->p2p_recs) return fail(ctx, MFKC_E_STATE, "mfkc_p2p_attach before mfkc_p2p_stage_create");
        ctx->p2p_peer_recs[rank] = ctx->p2p_recs; ctx->p2p_peer_cursor[rank] = ctx->p2p_cursor; ctx->p2p_ipc[rank] = false;
        return MFKC_OK;
    }
    if (!handles) return fail(ctx, MFKC_E_BADARG, "null argument");
    cudaIpcMemHandle_t h; void *p = nullptr;
    memcpy(&h, handles, 64);
    CU_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess)); ctx->p2p_peer_recs[rank] = reinterpret_cast<const uint4 *>(p);
    memcpy(&h, handles + 64, 64);
    CU_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess)); ctx->p2p_peer_cursor[rank] = reinterpret_cast<const unsigned int *>(p);
    ctx->p2p_ipc[rank] = true;
    return MFKC_OK;
}

// contexts of the same process (logical shards on one GPU in the tests, or one process driving several GPUs)
extern "C" int mfkc_p2p_attach_ctx(mfkc_ctx *ctx, uint32_t rank, mfkc_ctx *peer) {
    TRY(p2p_check(ctx));
    if (!peer || !peer->p2p_recs || rank >= (uint32_t)std::max(1, ctx->cfg.n_shards)) return fail(ctx, MFKC_E_BADARG, "bad peer");
    if (peer->p2p_seg_cap != ctx->p2p_seg_cap || peer->p2p_log2 != ctx->p2p_log2 || peer->p2p_bins != ctx->p2p_bins || peer->p2p_B != ctx->p2p_B ||
        peer->p2p_ovf_cap != ctx->p2p_ovf_cap) return fail(ctx, MFKC_E_BADARG, "peer staging geometry differs");
    if (peer->device != ctx->device) {
        CU_TRY(cudaSetDevice(ctx->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return fail(ctx, MFKC_E_CUDA, "peer access between the two devices is not available"); }
        cudaGetLastError();
    }
    ctx->p2p_peer_recs[rank] = peer->p2p_recs; ctx->p2p_peer_cursor[rank] = peer->p2p_cursor; ctx->p2p_ipc[rank] = false;
    return MFKC_OK;
}

extern "C" int mfkc_p2p_stage_reset(mfkc_ctx *ctx) {
    TRY(p2p_check(ctx));
    if (!ctx->p2p_recs) return MFKC_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaMemsetAsync(ctx->p2p_cursor, 0, ctx->p2p_n_cursor * sizeof(unsigned int), ctx->compute));
    CU_TRY(cudaMemsetAsync(ctx->p2p_kc, 0, P2P_MAX_PEERS * sizeof(unsigned long long), ctx->compute));
    if (ctx->p2p_bins) {
        const OvfLayout o = ovf_layout(ctx->p2p_recs + (uint64_t)std::max(1, ctx->cfg.n_shards) * ctx->p2p_B * ctx->p2p_seg_cap, ctx->p2p_ovf_cap);
        CU_TRY(cudaMemsetAsync(o.tags, 0xFF, (size_t)o.n_chunks * sizeof(unsigned long long), ctx->compute));
    }
    ctx->p2p_kmers_in = 0;
    if (ctx->p2p_bins) { ctx->os_counted = false; if (ctx->mode == 1) ctx->mode = 0; }
    return MFKC_OK;
}

static int p2p_extract_batch(mfkc_ctx *ctx, Staging &s, const uint8_t *d_bases, const uint64_t *d_offsets, uint32_t n_reads, uint64_t n_bases) {
    const int k = ctx->cfg.k;
    TRY(launch_mark(ctx, s, d_offsets, n_reads, n_bases, ctx->cfg.min_seq_len, 1, ctx->d_ctr));
    if (n_bases >= (uint64_t)k && ctx->p2p_bins) {
        ProfScope ps(ctx, P_EXTRACT_BUCKET, ctx->compute);
        extract_skm_kernel<4, TabAoS><<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
            d_bases, n_bases, s.d_flags, k, bin_stage(ctx), TabAoS{nullptr, 0}, ctx->d_ctr, ctx->p2p_kc);
    } else if (n_bases >= (uint64_t)k) {
        SkmStage st{};
        st.recs = ctx->p2p_recs; st.cursor = ctx->p2p_cursor; st.seg_cap = ctx->p2p_seg_cap;
        st.n_regions = (uint32_t)std::max(1, ctx->cfg.n_shards); st.region_shift = ctx->p2p_log2; st.win = 0;
        ProfScope ps(ctx, P_EXTRACT_BUCKET, ctx->compute);
        if (ctx->k128) {
            SkmStage128 s8; s8.recs = st.recs; s8.cursor = st.cursor; s8.seg_cap = st.seg_cap; s8.n_regions = st.n_regions; s8.region_shift = st.region_shift;
            extract_skm128_kernel<2><<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
                d_bases, n_bases, s.d_flags, k, s8, nullptr, 0, ctx->d_ctr, ctx->p2p_kc);
        } else
        extract_skm_kernel<2, TabAoS><<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
            d_bases, n_bases, s.d_flags, k, st, TabAoS{nullptr, 0}, ctx->d_ctr, ctx->p2p_kc);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(s.ev_done, ctx->compute));
    ctx->dirty = true; ctx->hist_valid = false; ctx->em_valid = false;
    return MFKC_OK;
}

extern "C" int mfkc_p2p_extract(mfkc_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offsets, uint32_t n_reads, uint64_t n_bases) {
    TRY(p2p_check(ctx));
    if (!d_bases || !d_offsets) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (!ctx->p2p_recs) return fail(ctx, MFKC_E_STATE, "mfkc_p2p_extract before mfkc_p2p_stage_create");
    if (n_reads == 0) return MFKC_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    Staging &s = ctx->st[ctx->next_buf];
    ctx->next_buf = (ctx->next_buf + 1) % N_STAGE;
    TRY(ensure_staging(ctx, s, n_bases, 0, false));
    return p2p_extract_batch(ctx, s, d_bases, d_offsets, n_reads, n_bases);
}

// host buffers: same copy pipeline as mfkc_submit_reads (N_STAGE batches in flight), extraction into the p2p staging
extern "C" int mfkc_p2p_submit_reads(mfkc_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads) {
    TRY(p2p_check(ctx));
    if (!ctx->p2p_recs) return fail(ctx, MFKC_E_STATE, "mfkc_p2p_submit_reads before mfkc_p2p_stage_create");
    return submit_host(ctx, bases, offsets, n_reads, true);
}

extern "C" int mfkc_p2p_counts(mfkc_ctx *ctx, uint64_t *kmers_per_owner) {
    TRY(p2p_check(ctx));
    if (!kmers_per_owner || !ctx->p2p_kc) return fail(ctx, MFKC_E_BADARG, "null argument");
    CU_TRY(cudaSetDevice(ctx->device));
    const int ns = std::max(1, ctx->cfg.n_shards);
    if (ctx->p2p_bins) TRY(place_overflow(ctx));            // the peers find this rank's overflow records through the chunk pool
    CU_TRY(cudaMemcpyAsync(ctx->h_bucket, ctx->p2p_kc, ns * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->compute));
    CU_TRY(cudaStreamSynchronize(ctx->compute));          // every record of this rank is in its staging buffer
    for (int i = 0; i < ns; i++) kmers_per_owner[i] = ctx->h_bucket[i];
    return MFKC_OK;
}

extern "C" int mfkc_p2p_drain(mfkc_ctx *ctx, uint64_t n_kmers_in) {
    TRY(p2p_check(ctx));
    const uint32_t ns = (uint32_t)std::max(1, ctx->cfg.n_shards);
    for (uint32_t i = 0; i < ns; i++) if (!ctx->p2p_peer_recs[i]) return fail(ctx, MFKC_E_STATE, "mfkc_p2p_drain: not every peer is attached");
    if (ctx->p2p_bins) {
        // bin-local: nothing is counted yet -- the result calls (stats / histogram / emit_begin) run bin_count_kernel over the
        // peers' staging buffers, which therefore must stay untouched until every rank has fetched its results
        ctx->p2p_kmers_in = n_kmers_in; ctx->mode = 1; ctx->sample_open = true; ctx->os_counted = false;
        ctx->dirty = true; ctx->hist_valid = false; ctx->em_valid = false;
        return MFKC_OK;
    }
    if (n_kmers_in == 0) return MFKC_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(ensure_table_mode(ctx));
    P2PPeers pp;
    for (uint32_t i = 0; i < (uint32_t)P2P_MAX_PEERS; i++) { pp.recs[i] = i < ns ? ctx->p2p_peer_recs[i] : nullptr; pp.cursor[i] = i < ns ? ctx->p2p_peer_cursor[i] : nullptr; }
    pp.seg_cap = ctx->p2p_seg_cap; pp.n_peers = ns; pp.me = (uint32_t)ctx->cfg.shard_id; pp.log2_buckets = ctx->p2p_log2;
    const uint64_t n_buckets = 1ull << ctx->p2p_log2;
    const uint64_t recs_per_bucket = n_kmers_in / 4 / n_buckets + 1;
    static const int env_bpb = getenv("MFKC_P2P_BPB") ? atoi(getenv("MFKC_P2P_BPB")) : 0;
    uint32_t bpb = (uint32_t)std::min<uint64_t>(592, std::max<uint64_t>(1, recs_per_bucket / 512));
    if (env_bpb > 0) bpb = (uint32_t)env_bpb;
    // The buckets are drained in a few chunks.  A chunk is a fixed fraction of the minimizer-hash space, so it holds
    // that fraction of the incoming k-mers (thousands of minimizers per chunk: +-1 %, covered by the 10 % margin and
    // the 8 % of the table the load limit keeps free).  reserve_slots() thereby sees the exact distinct count of the
    // chunks before (it synchronises only when its bound is exceeded) and a right-sized table is not grown on the
    // strength of "every incoming k-mer could be new".  The records are geometry-free, so growth in between is fine.
    static const int env_chunks = getenv("MFKC_P2P_CHUNKS") ? atoi(getenv("MFKC_P2P_CHUNKS")) : 0;
    // chunk size: 15 % of the table, so that "distinct so far + chunk" stays below the growth threshold of a table that
    // was sized for load 0.4
    uint64_t chunks = env_chunks > 0 ? (uint64_t)env_chunks : (uint64_t)((double)n_kmers_in * 1.1 / (0.15 * (double)ctx->cap)) + 1;
    if ((double)(ctx->distinct_ub + n_kmers_in) <= kMaxLoad * (double)ctx->cap) chunks = 1;      // fits as a whole
    chunks = std::min<uint64_t>(chunks, n_buckets);
    for (uint64_t c = 0; c < chunks; c++) {
        const uint64_t b0 = n_buckets * c / chunks, b1 = n_buckets * (c + 1) / chunks;
        const uint64_t share = chunks == 1 ? n_kmers_in : (uint64_t)((double)n_kmers_in * (double)(b1 - b0) / (double)n_buckets * 1.1) + 1024;
        TRY(reserve_slots(ctx, share));                   // may drain local staging and grow the table
        ctx->kmers_ub_total += share; ctx->recv_since_base += share;
        {
            ProfScope ps(ctx, P_DRAIN, ctx->compute);
            if (ctx->k128) {
                if (!ctx->d_peer_recs) {
                    CU_TRY(cudaMalloc(&ctx->d_peer_recs, P2P_MAX_PEERS * sizeof(void *)));
                    CU_TRY(cudaMalloc(&ctx->d_peer_cursor, P2P_MAX_PEERS * sizeof(void *)));
                }
                CU_TRY(cudaMemcpyAsync(ctx->d_peer_recs, pp.recs, P2P_MAX_PEERS * sizeof(void *), cudaMemcpyHostToDevice, ctx->compute));
                CU_TRY(cudaMemcpyAsync(ctx->d_peer_cursor, pp.cursor, P2P_MAX_PEERS * sizeof(void *), cudaMemcpyHostToDevice, ctx->compute));
                CU_TRY(cudaStreamSynchronize(ctx->compute));              // pp lives on this stack frame
                drain_p2p128_kernel<<<(unsigned)((b1 - b0) * bpb), 256, 0, ctx->compute>>>(
                    ctx->d_peer_recs, ctx->d_peer_cursor, ns, pp.me, pp.log2_buckets, pp.seg_cap, (uint32_t)b0, bpb, ctx->cfg.k,
                    ctx->tab128, ctx->cap, ctx->n_regions, ctx->region_shift, ctx->d_ctr);
            } else
            drain_p2p_kernel<TabAoS><<<(unsigned)((b1 - b0) * bpb), 256, 0, ctx->compute>>>(
                pp, (uint32_t)b0, bpb, ctx->cfg.k, tab_aos(ctx), ctx->n_regions, ctx->region_shift, table_win(ctx), ctx->d_ctr);
        }
        CU_TRY(cudaGetLastError());
    }
    CU_TRY(cudaMemcpyAsync(ctx->h_drain_snap, &ctx->d_ctr->distinct, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->compute));
    CU_TRY(cudaEventRecord(ctx->ev_drain, ctx->compute));
    ctx->drain_pending = true; ctx->kmers_at_drain = ctx->kmers_ub_total;
    ctx->dirty = true; ctx->hist_valid = false; ctx->em_valid = false;
    return MFKC_OK;
}

// ------------------------------------------------------------------------------------------
// features-calculator
// ------------------------------------------------------------------------------------------
extern "C" int mfkc_fc_load_components(mfkc_ctx *ctx, const int64_t *keys, const uint64_t *comp_offsets, uint32_t n_comp) {
    if (!ctx || !comp_offsets || (n_comp && comp_offsets[n_comp] && !keys)) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (ctx->k128) return fail(ctx, MFKC_E_BADARG, "features-calculator works on 64-bit keys (k <= 31), like the reference");
    CU_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->compute;
    cudaFree(ctx->fc_tab); cudaFree(ctx->fc_keys); cudaFree(ctx->fc_off); cudaFree(ctx->fc_bloom);
    ctx->fc_tab = nullptr; ctx->fc_keys = nullptr; ctx->fc_off = nullptr; ctx->fc_bloom = nullptr; ctx->fc_bmask = 0;
    const uint64_t nk = comp_offsets[n_comp] - comp_offsets[0];
    ctx->fc_nkeys = nk; ctx->fc_ncomp = n_comp;
    ctx->fc_cap = std::max<uint64_t>(nk * 2 + 64, 1024);
    CU_TRY(cudaMalloc(&ctx->fc_tab, ctx->fc_cap * sizeof(FcSlot)));
    if (!getenv("MFKC_FC_NO_BLOOM")) {
        uint64_t bits = 1ull << 16;
        while (bits < 8 * nk && bits < (1ull << 29)) bits <<= 1;            // 8-16 bits per key, 64 MiB at most (stays in L2)
        CU_TRY(cudaMalloc(&ctx->fc_bloom, bits / 8));
        CU_TRY(cudaMemsetAsync(ctx->fc_bloom, 0, bits / 8, st));
        ctx->fc_bmask = (uint32_t)(bits - 1);
    }
    CU_TRY(cudaMalloc(&ctx->fc_keys, std::max<uint64_t>(nk, 1) * 8));
    CU_TRY(cudaMalloc(&ctx->fc_off, ((size_t)n_comp + 1) * 8));
    std::vector<uint64_t> off(n_comp + 1);
    for (uint32_t i = 0; i <= n_comp; i++) off[i] = comp_offsets[i] - comp_offsets[0];
    CU_TRY(cudaMemcpyAsync(ctx->fc_off, off.data(), ((size_t)n_comp + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nk) CU_TRY(cudaMemcpyAsync(ctx->fc_keys, keys + comp_offsets[0], nk * 8, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(ctx->d_fc_ctr, 0, sizeof(Counters), st));
    fc_clear_kernel<<<grid_for(ctx, ctx->fc_cap, 256, 8), 256, 0, st>>>(ctx->fc_tab, ctx->fc_cap, 1);
    if (nk) {
        ProfScope ps(ctx, P_FC_BUILD, st);
        fc_build_kernel<<<grid_for(ctx, nk, 256, 8), 256, 0, st>>>(ctx->fc_keys, nk, ctx->fc_tab, ctx->fc_cap, ctx->d_fc_ctr, ctx->fc_bloom, ctx->fc_bmask);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(st));
    return MFKC_OK;
}

extern "C" int mfkc_fc_reset_values(mfkc_ctx *ctx) {
    if (!ctx) return MFKC_E_BADARG;
    if (!ctx->fc_tab) return fail(ctx, MFKC_E_STATE, "no components loaded");
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(sync_all(ctx));
    fc_clear_kernel<<<grid_for(ctx, ctx->fc_cap, 256, 8), 256, 0, ctx->compute>>>(ctx->fc_tab, ctx->fc_cap, 0);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(ctx->compute));
    return MFKC_OK;
}

// stage host records on the device (through the bases staging buffer of slot 0)
static int stage_records(mfkc_ctx *ctx, Staging &s, const uint8_t *recs, uint64_t n) {
    TRY(ensure_staging(ctx, s, n * 10, 1, true));
    CU_TRY(cudaMemcpyAsync(s.d_bases, recs, (size_t)n * 10, cudaMemcpyHostToDevice, s.stream));
    return MFKC_OK;
}

extern "C" int mfkc_fc_set_selected(mfkc_ctx *ctx, const uint8_t *be_records, uint64_t n_records) {
    if (!ctx) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(sync_all(ctx));
    cudaStream_t st = ctx->compute;
    if (!be_records) {                       // no --selected: every component k-mer is considered
        cudaFree(ctx->fc_sel); ctx->fc_sel = nullptr; ctx->fc_sel_cap = 0; ctx->fc_sel_n = 0;
        return MFKC_OK;
    }
    // (re)size the selected set for the cumulative number of records
    const uint64_t need = (ctx->fc_sel_n + n_records) * 2 + 1024;
    if (need > ctx->fc_sel_cap) {
        Slot *nt = nullptr;
        TRY(table_alloc(ctx, need, &nt, true));
        if (ctx->fc_sel) {
            TableGeom pg; pg.cap = need; pg.n_regions = 1; pg.region_shift = 0; pg.k = ctx->cfg.k; pg.minimizer = 0; pg.win = 0;
            rehash_kernel<<<grid_for(ctx, ctx->fc_sel_cap, 256, 8), 256, 0, st>>>(ctx->fc_sel, ctx->fc_sel_cap, nt, pg);
            CU_TRY(cudaStreamSynchronize(st));
            cudaFree(ctx->fc_sel);
        }
        ctx->fc_sel = nt; ctx->fc_sel_cap = need;
    }
    if (n_records) {
        TRY(stage_records(ctx, ctx->st[0], be_records, n_records));
        CU_TRY(cudaStreamSynchronize(ctx->st[0].stream));
        fc_selected_kernel<<<grid_for(ctx, n_records, 256, 8), 256, 0, st>>>(ctx->st[0].d_bases, n_records, ctx->fc_sel,
                                                                            ctx->fc_sel_cap, ctx->d_fc_ctr);
        CU_TRY(cudaGetLastError());
    }
    CU_TRY(cudaStreamSynchronize(st));
    ctx->fc_sel_n += n_records;
    return MFKC_OK;
}

extern "C" int mfkc_fc_add_records(mfkc_ctx *ctx, const uint8_t *be_records, uint64_t n_records) {
    if (!ctx || (!be_records && n_records)) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (!ctx->fc_tab) return fail(ctx, MFKC_E_STATE, "no components loaded");
    if (n_records == 0) return MFKC_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    Staging &s = ctx->st[ctx->next_buf];
    ctx->next_buf = (ctx->next_buf + 1) % N_STAGE;
    CU_TRY(cudaStreamWaitEvent(s.stream, s.ev_done, 0));
    TRY(stage_records(ctx, s, be_records, n_records));
    CU_TRY(cudaEventRecord(s.ev_copy, s.stream));
    CU_TRY(cudaStreamWaitEvent(ctx->compute, s.ev_copy, 0));
    {
        ProfScope ps(ctx, P_FC_RECORDS, ctx->compute);
        fc_records_kernel<<<grid_for(ctx, n_records, 256, 8), 256, 0, ctx->compute>>>(s.d_bases, n_records, ctx->fc_tab, ctx->fc_cap, ctx->fc_bloom, ctx->fc_bmask);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(s.ev_done, ctx->compute));
    CU_TRY(cudaEventSynchronize(s.ev_copy));
    return MFKC_OK;
}

// records that are still on the device: the emit arrays of a counter context on the same GPU (after mfkc_emit_begin)
extern "C" int mfkc_fc_add_emitted(mfkc_ctx *ctx, mfkc_ctx *counter) {
    if (!ctx || !counter) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (!ctx->fc_tab) return fail(ctx, MFKC_E_STATE, "no components loaded");
    if (!counter->em_valid || counter->k128) return fail(ctx, MFKC_E_STATE, "mfkc_fc_add_emitted needs a counter after mfkc_emit_begin (k <= 31)");
    if (counter->device != ctx->device) return fail(ctx, MFKC_E_BADARG, "the counter lives on another device");
    if (counter->em_n == 0) return MFKC_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    {   // the counter's emit finished with a host synchronisation, so its arrays are complete
        ProfScope ps(ctx, P_FC_RECORDS, ctx->compute);
        fc_pairs_kernel<<<grid_for(ctx, counter->em_n, 256, 8), 256, 0, ctx->compute>>>(counter->em_keys, counter->em_counts, counter->em_n, ctx->fc_tab, ctx->fc_cap, ctx->fc_bloom, ctx->fc_bmask);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(ctx->compute));             // the counter may reset its arrays as soon as this returns
    return MFKC_OK;
}

extern "C" int mfkc_fc_add_reads(mfkc_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads) {
    if (!ctx || (!bases && n_reads) || !offsets) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (!ctx->fc_tab) return fail(ctx, MFKC_E_STATE, "no components loaded");
    if (n_reads == 0) return MFKC_OK;
    CU_TRY(cudaSetDevice(ctx->device));
    const uint64_t base0 = offsets[0];
    const uint64_t n_bases = offsets[n_reads] - base0;
    Staging &s = ctx->st[ctx->next_buf];
    ctx->next_buf = (ctx->next_buf + 1) % N_STAGE;
    TRY(ensure_staging(ctx, s, n_bases, (uint64_t)n_reads + 1, true));
    CU_TRY(cudaStreamWaitEvent(s.stream, s.ev_done, 0));
    CU_TRY(cudaMemcpyAsync(s.d_bases, bases + base0, n_bases, cudaMemcpyHostToDevice, s.stream));
    CU_TRY(cudaMemcpyAsync(s.d_offsets, offsets, ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, s.stream));
    CU_TRY(cudaEventRecord(s.ev_copy, s.stream));
    CU_TRY(cudaStreamWaitEvent(ctx->compute, s.ev_copy, 0));
    // ReadsPresenceWorker has no minSeqLen (src/io/IOUtils.java:806-825)
    TRY(launch_mark(ctx, s, s.d_offsets, n_reads, n_bases, 0, 0, ctx->d_fc_ctr));
    if (n_bases >= (uint64_t)ctx->cfg.k) {
        ProfScope ps(ctx, P_FC_READS, ctx->compute);
        SinkPresence sink{ctx->fc_tab, ctx->fc_cap, ctx->fc_bloom, ctx->fc_bmask};
        extract_kernel<SinkPresence><<<extract_grid(ctx, n_bases), EX_THREADS, 0, ctx->compute>>>(
            s.d_bases, n_bases, s.d_flags, ctx->cfg.k, sink, ctx->d_fc_ctr);
        CU_TRY(cudaGetLastError());
    }
    CU_TRY(cudaEventRecord(s.ev_done, ctx->compute));
    CU_TRY(cudaEventSynchronize(s.ev_copy));
    return MFKC_OK;
}

extern "C" int mfkc_fc_features(mfkc_ctx *ctx, int64_t threshold, int64_t *vec, uint64_t *found, uint64_t *cnt) {
    if (!ctx || !vec || !found || !cnt) return fail(ctx, MFKC_E_BADARG, "null argument");
    if (!ctx->fc_tab) return fail(ctx, MFKC_E_STATE, "no components loaded");
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(sync_all(ctx));
    CU_TRY(cudaMemcpyAsync(ctx->h_ctr, ctx->d_fc_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->compute));
    CU_TRY(cudaStreamSynchronize(ctx->compute));
    if (ctx->h_ctr->bad_chars) return fail(ctx, MFKC_E_FORMAT, "Incorrect nucleotide char in submitted reads (only AaCcGgTt are accepted)");
    const uint32_t nc = ctx->fc_ncomp;
    if (nc == 0) return MFKC_OK;
    cudaStream_t st = ctx->compute;
    long long *d_vec = nullptr; unsigned long long *d_found = nullptr, *d_cnt = nullptr;
    TMP_ALLOC(d_vec, (size_t)nc * 8); TMP_ALLOC(d_found, (size_t)nc * 8); TMP_ALLOC(d_cnt, (size_t)nc * 8);
    {
        ProfScope ps(ctx, P_FC_FEATURES, st);
        const int grid = grid_for(ctx, (uint64_t)nc * 32, 256, 8);
        fc_features_kernel<<<grid, 256, 0, st>>>(ctx->fc_keys, ctx->fc_off, nc, ctx->fc_tab, ctx->fc_cap, ctx->fc_sel,
                                                ctx->fc_sel_cap, (long long)threshold, d_vec, d_found, d_cnt);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpyAsync(vec, d_vec, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(found, d_found, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(cnt, d_cnt, (size_t)nc * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    TMP_FREE(d_vec); TMP_FREE(d_found); TMP_FREE(d_cnt);
    return MFKC_OK;
}

// ------------------------------------------------------------------------------------------
// synthetic reads on the device
// ------------------------------------------------------------------------------------------
extern "C" int mfkc_synth_tables_build(const mfkc_synth_cfg *cfg, mfkc_synth_tables *t);   // host_io.cpp

__global__ void __launch_bounds__(256)
synth_kernel(const mfkc_synth_tables *__restrict__ t, const uint64_t *__restrict__ kept_index, uint64_t n_kept,
             uint8_t *__restrict__ out, uint64_t *__restrict__ offsets) {
    const uint32_t L = t->read_len;
    const uint64_t total = n_kept * L;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = i / L;
        const uint32_t b = (uint32_t)(i - j * L);
        mfkc_synth_read r;
        mfkc_synth_read_header(*t, kept_index[j], r);
        out[i] = mfkc_synth_read_base(*t, r, b);
        if (b == 0) offsets[j] = i;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) offsets[n_kept] = total;
}

extern "C" int mfkc_synth_reads_device(mfkc_ctx *ctx, const mfkc_synth_cfg *cfg, uint64_t first_read, uint64_t n_reads,
                                       uint8_t *d_bases, uint64_t *d_offsets, uint64_t *n_kept) {
    if (!ctx || !cfg || !d_bases || !d_offsets || !n_kept) return fail(ctx, MFKC_E_BADARG, "null argument");
    CU_TRY(cudaSetDevice(ctx->device));
    mfkc_synth_tables *t = new mfkc_synth_tables();
    int r = mfkc_synth_tables_build(cfg, t);
    if (r != MFKC_OK) { delete t; return fail(ctx, r, "bad synthetic-read configuration"); }
    std::vector<uint64_t> kept; kept.reserve(n_reads);
    for (uint64_t i = 0; i < n_reads; i++) if (!mfkc_synth_read_has_n(*t, first_read + i)) kept.push_back(first_read + i);
    cudaStream_t st = ctx->compute;
    if (!ctx->d_synth) CU_TRY(cudaMalloc(&ctx->d_synth, sizeof(mfkc_synth_tables)));
    uint64_t *d_idx = nullptr;
    CU_TRY(cudaMalloc(&d_idx, std::max<size_t>(kept.size(), 1) * 8));
    CU_TRY(cudaMemcpyAsync(ctx->d_synth, t, sizeof(mfkc_synth_tables), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_idx, kept.data(), kept.size() * 8, cudaMemcpyHostToDevice, st));
    {
        ProfScope ps(ctx, P_SYNTH, st);
        synth_kernel<<<grid_for(ctx, kept.size() * (uint64_t)t->read_len + 1, 256, 8), 256, 0, st>>>(
            ctx->d_synth, d_idx, kept.size(), d_bases, d_offsets);
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(st));
    cudaFree(d_idx);
    *n_kept = kept.size();
    delete t;
    return MFKC_OK;
}

// ------------------------------------------------------------------------------------------
// raw device helpers, timing, profile, GUPS
// ------------------------------------------------------------------------------------------
extern "C" int mfkc_device_alloc(mfkc_ctx *ctx, size_t bytes, void **d_ptr) {
    if (!ctx || !d_ptr) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return MFKC_OK;
}
extern "C" int mfkc_device_free(mfkc_ctx *ctx, void *d_ptr) {
    if (!ctx) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaFree(d_ptr));
    return MFKC_OK;
}
extern "C" int mfkc_memcpy_h2d(mfkc_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
    if (!ctx) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    // On the compute stream and awaited: a legacy-stream cudaMemcpy from pageable memory may return
    // while its last staged chunk is still in flight, and the (non-blocking) compute stream does not
    // wait for the legacy stream -- a kernel launched right after could read the old bytes.
    CU_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->compute));
    CU_TRY(cudaStreamSynchronize(ctx->compute));
    return MFKC_OK;
}
extern "C" int mfkc_memcpy_d2h(mfkc_ctx *ctx, void *h_dst, const void *d_src, size_t bytes) {
    if (!ctx) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->compute));   // after the compute stream's kernels
    CU_TRY(cudaStreamSynchronize(ctx->compute));
    return MFKC_OK;
}
extern "C" int mfkc_device_sync(mfkc_ctx *ctx) {
    if (!ctx) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaDeviceSynchronize());
    return MFKC_OK;
}

// Every kernel runs on the compute stream and every H2D copy is awaited by it, so two events
// on the compute stream bracket all device work submitted in between.
extern "C" int mfkc_timer_start(mfkc_ctx *ctx) {
    if (!ctx) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    TRY(sync_all(ctx));
    CU_TRY(cudaEventRecord(ctx->t0, ctx->compute));
    return MFKC_OK;
}
extern "C" int mfkc_timer_stop_ms(mfkc_ctx *ctx, float *ms) {
    if (!ctx || !ms) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    CU_TRY(cudaEventRecord(ctx->t1, ctx->compute));
    CU_TRY(cudaEventSynchronize(ctx->t1));
    CU_TRY(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return MFKC_OK;
}

extern "C" int mfkc_profile_enable(mfkc_ctx *ctx, int on) { if (!ctx) return MFKC_E_BADARG; ctx->profiling = on != 0; return MFKC_OK; }
extern "C" int mfkc_profile_reset(mfkc_ctx *ctx) {
    if (!ctx) return MFKC_E_BADARG;
    drain_timings(ctx);
    for (int i = 0; i < P_NSLOTS; i++) { ctx->prof_ms[i] = 0; ctx->prof_launches[i] = 0; }
    return MFKC_OK;
}
extern "C" int mfkc_profile_get(mfkc_ctx *ctx, int slot, double *ms, uint64_t *launches) {
    if (!ctx || slot < 0 || slot >= P_NSLOTS) return MFKC_E_BADARG;
    drain_timings(ctx);
    if (ms) *ms = ctx->prof_ms[slot];
    if (launches) *launches = ctx->prof_launches[slot];
    return MFKC_OK;
}
extern "C" const char *mfkc_profile_name(int slot) { return slot >= 0 && slot < P_NSLOTS ? kProfNames[slot] : nullptr; }

extern "C" int mfkc_gups_ex(mfkc_ctx *ctx, uint64_t bytes, uint64_t n_updates, int mode, uint64_t window_bytes,
                            uint32_t blocks_per_window, float *ms) {
    if (!ctx || !ms || bytes < 32 || mode < 0 || mode > 5) return MFKC_E_BADARG;
    CU_TRY(cudaSetDevice(ctx->device));
    unsigned long long *tab = nullptr;
    CU_TRY(cudaMalloc(&tab, bytes));
    cudaStream_t st = ctx->compute;
    CU_TRY(cudaMemsetAsync(tab, 0, bytes, st));
    const uint64_t n_sectors = bytes / 32;
    uint64_t win = window_bytes / 32;
    if (win == 0 || win > n_sectors) win = n_sectors;
    if (blocks_per_window == 0) blocks_per_window = 1;
    int grid = ctx->sm_count * 8;
    if (mode >= 3) {
        grid = (int)std::min<uint64_t>((n_sectors + win - 1) / win * blocks_per_window, 1u << 30);
        if (grid < (int)blocks_per_window) grid = blocks_per_window;
        grid -= grid % blocks_per_window;
    }
    cudaEvent_t a = get_event(ctx), b = get_event(ctx);
    gups_kernel<<<grid, 256, 0, st>>>(tab, n_sectors, n_updates / 8 + 1, 1, mode, win, blocks_per_window);   // warm-up
    CU_TRY(cudaEventRecord(a, st));
    {
        ProfScope ps(ctx, P_GUPS, st);
        gups_kernel<<<grid, 256, 0, st>>>(tab, n_sectors, n_updates, 0x9e3779b9ULL, mode, win, blocks_per_window);
    }
    CU_TRY(cudaEventRecord(b, st));
    CU_TRY(cudaEventSynchronize(b));
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventElapsedTime(ms, a, b));
    ctx->ev_pool.push_back(a); ctx->ev_pool.push_back(b);
    CU_TRY(cudaFree(tab));
    return MFKC_OK;
}

extern "C" int mfkc_gups(mfkc_ctx *ctx, uint64_t bytes, uint64_t n_updates, float *ms) {
    return mfkc_gups_ex(ctx, bytes, n_updates, 1, 0, 0, ms);
}

#include "kset_api.inl"
