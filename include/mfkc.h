/*
 * mfkc.h -- C ABI of libmfkc: the B200-native (sm_100a) k-mer counting hot path
 * of MetaFast (kmer-counter-many -> .kmers.bin / .stat.txt -> features-calculator).
 *
 * Plain C: opaque handles, pointers and sizes only; no callbacks, no structs by
 * value, no C++ exceptions across the boundary.  Callable from Panama FFM
 * (Linker.downcallHandle), JNI, ctypes or C++ alike (INTEGRATION.md shows the
 * Java binding a MetaFast maintainer would add).
 *
 * Every entry point names the reference interface it replaces.  Prefixes:
 *   src/    = ctlab/metafast  src/
 *   [itmo]/ = lib/itmo-assembler-src.jar!/ru/ifmo/genetics/
 *
 * Conventions
 *   - return 0 (MFKC_OK) or a negative MFKC_E_* code; text via mfkc_last_error().
 *   - the caller owns every buffer it passes; the library owns device memory
 *     and the pinned buffers it handed out.
 *   - a context is driven by one host thread at a time; several contexts may
 *     coexist (one per GPU / per sample stream).
 *   - there is NO CPU fallback: without a CUDA device mfkc_create fails with
 *     MFKC_E_CUDA.
 */
#ifndef MFKC_H
#define MFKC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFKC_ABI_VERSION 1

enum {
    MFKC_OK = 0,
    MFKC_E_BADARG = -1,     /* IllegalArgumentException / System.exit(1) paths of the tools */
    MFKC_E_CUDA = -2,
    MFKC_E_NCCL = -3,
    MFKC_E_TABLE_FULL = -4, /* table cannot grow any further (device memory exhausted) */
    MFKC_E_OOM = -5,
    MFKC_E_STATE = -6,      /* call sequence violated (e.g. emit_next before emit_begin) */
    MFKC_E_IO = -7,         /* file cannot be opened / read / written */
    MFKC_E_FORMAT = -8      /* malformed FASTA/FASTQ/.kmers.bin/components.bin, bad nucleotide */
};

/* HASH        = open-addressing hash counting, partitioned by minimizer (default).  k <= 31 on one GPU: the
 *               sample's super-k-mer records are staged per minimizer BIN and every bin is counted by one
 *               CTA in a SHARED-MEMORY table (records streamed in with TMA bulk copies), which emits
 *               histogram + filtered (key,count) pairs directly; only bins that overflow go through the
 *               HBM-resident table.  Samples that outgrow the staging buffer, k > 31 and sharded contexts
 *               use HASH_TABLE's path;
 * HASH_TABLE  = HBM-resident open-addressing table, region-blocked: keys are partitioned by table region
 *               and drained region by region so that every upsert hits L2 (round 1's default);
 * SORT        = accumulate keys, radix sort, run-length encode;
 * HASH_DIRECT = the same table updated straight from the extraction kernel (one random DRAM
 *               sector per k-mer) -- kept as the measured baseline of the partitioned designs. */
enum { MFKC_VARIANT_HASH = 0, MFKC_VARIANT_SORT = 1, MFKC_VARIANT_HASH_DIRECT = 2, MFKC_VARIANT_HASH_TABLE = 3 };

#define MFKC_MAX_COUNT 32767        /* Short.MAX_VALUE: [itmo]/utils/NumUtils.java:21-26 */
#define MFKC_HIST_BINS 32768        /* histogram index = count, 1..32767 */

typedef struct mfkc_ctx mfkc_ctx;

/* Configuration; zero-initialise, set struct_size = sizeof(mfkc_cfg). */
typedef struct mfkc_cfg {
    uint32_t struct_size;
    int32_t  k;                 /* 1..31: 64-bit keys (reference range, src/tools/KmersCounterMain.java:66-73);
                                   32..63: 128-bit keys (extension, no reference behaviour) */
    int32_t  min_seq_len;       /* minSeqLen of IOUtils.loadReads (src/io/IOUtils.java:761); 0 for the counter */
    int32_t  device;            /* CUDA device ordinal */
    int32_t  variant;           /* MFKC_VARIANT_HASH | MFKC_VARIANT_SORT | MFKC_VARIANT_HASH_DIRECT | MFKC_VARIANT_HASH_TABLE */
    int32_t  n_shards;          /* hash-range sharding: number of key-space shards (GPUs); 0/1 = unsharded */
    int32_t  shard_id;          /* the shard this context owns */
    int32_t  reserved0;
    uint64_t table_slots;       /* initial table capacity in slots; 0 = derive from expected_distinct / default.  Setting
                                   table_slots, staging_bytes or region_shift pins the table geometry (HASH_TABLE's path) */
    uint64_t expected_distinct; /* sizing hint (distinct k-mers); 0 = unknown, table grows x2 on demand */
    uint64_t max_table_bytes;   /* growth limit; 0 = 80 % of free device memory */
    uint64_t staging_bytes;     /* HASH: key staging buffer; 0 = adaptive (starts at 4 batches, doubles when full) */
    uint32_t region_shift;      /* HASH_TABLE path: log2(table slots per region); 0 = 17 (2 MiB regions of 16-byte slots) */
    uint32_t reserved2;
    uint64_t expected_kmers;    /* sizing hint: k-mer INSTANCES of the sample (<= bases in the input files).  HASH: staging
                                   holds them all (one drain) and, without expected_distinct, the table is sized for the
                                   all-distinct worst case at load 0.85, memory permitting */
    uint64_t reserved1[1];
} mfkc_cfg;

/* ---- lifecycle ------------------------------------------------------------------------
 * Replaces `new BigLong2ShortHashMap(..)` + the ReadsLoadWorker pool set-up in
 * IOUtils.loadReads (src/io/IOUtils.java:772-781). */
int  mfkc_abi_version(void);
int  mfkc_device_count(void);                      /* 0 when no CUDA device / driver */
int  mfkc_create(const mfkc_cfg *cfg, mfkc_ctx **out);
void mfkc_destroy(mfkc_ctx *ctx);
const char *mfkc_last_error(const mfkc_ctx *ctx);  /* ctx may be NULL: last create() failure */
int  mfkc_reset(mfkc_ctx *ctx);                    /* next sample: empty table, keep allocations */

/* Pinned host buffers for the Java side to wrap as direct ByteBuffer / MemorySegment. */
int  mfkc_pinned_alloc(mfkc_ctx *ctx, size_t bytes, void **host_ptr);
int  mfkc_pinned_free(mfkc_ctx *ctx, void *host_ptr);

/* ---- ingest: replaces ReadsWorker.process(List<Dna>) (src/io/ReadsWorker.java:25,
 * src/io/IOUtils.java:756-769) fed by ReadsDispatcher.getWorkRange
 * (src/io/ReadsDispatcher.java:34-53) --------------------------------------------------
 * bases   = ASCII nucleotides (AaCcGgTt only) of reads that already passed the parser
 *           rules (N-drop, phred-0 drop: mfkc_reader_* below does that);
 * offsets = n_reads+1 byte offsets into `bases` (read i = bases[offsets[i]..offsets[i+1])).
 * Host-buffer version: copies asynchronously, returns once `bases`/`offsets` may be
 * refilled; counting continues in the background (double-buffered).
 * Any batch size is accepted (the reference uses 32768 reads, src/io/IOUtils.java:29). */
int  mfkc_submit_reads(mfkc_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads);
/* Same, for inputs already resident in device memory (no copy; used for kernel-only timing). */
int  mfkc_submit_reads_device(mfkc_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offsets,
                              uint32_t n_reads, uint64_t n_bases);
/* All submitted work is counted when this returns (latch.await(), src/io/IOUtils.java:853). */
int  mfkc_flush(mfkc_ctx *ctx);

/* ---- results: replaces IOUtils.printKmers (src/io/IOUtils.java:45-71) and the statistics
 * of loadReads (src/io/IOUtils.java:783-800) --------------------------------------------
 * stats[0] = distinct k-mers (hm.size()), [1] = k-mer instances counted,
 * [2] = totalSeq, [3] = goodSeq, [4] = totalLen, [5] = goodLen. */
int  mfkc_stats(mfkc_ctx *ctx, uint64_t stats[6]);
/* hist[c] = number of distinct k-mers with (saturated) count c, ALL entries
 * (QuickQuantitativeStatistics, src/io/IOUtils.java:59). */
int  mfkc_histogram(mfkc_ctx *ctx, uint64_t hist[MFKC_HIST_BINS]);
/* Diagnostics of MFKC_VARIANT_HASH's bin-local mode, valid after a result call (stats / histogram / emit_begin):
 * out[0] = 1 if the current sample is counted bin-locally, 0 if it uses (or fell back to) the region-blocked table;
 * [1] = bins, [2] = records per staging segment, [3] = heavy (bin, sub-range) entries counted through the table,
 * [4] = their records, [5] = passes that were split, [6] = records in the overflow list, [7] = records staged. */
int  mfkc_bin_stats(mfkc_ctx *ctx, uint64_t out[8]);
/* Select entries with count > threshold (src/io/IOUtils.java:61), order them by ascending key;
 * *n_good = how many.  Also (re)computes the histogram. */
int  mfkc_emit_begin(mfkc_ctx *ctx, int32_t threshold, uint64_t *n_good);
/* Next chunk of big-endian records (int64 key + int16 count = 10 bytes for k<=31; 16+2 bytes
 * for k>=32), a whole number of records; *written = 0 at the end. */
int  mfkc_emit_next(mfkc_ctx *ctx, uint8_t *out, size_t cap, size_t *written);
/* Device-resident result (for on-device consumers / kernel-only timing): pointers stay valid
 * until the next submit/reset/emit_begin.  keys ascending; counts saturated. */
int  mfkc_emit_device(mfkc_ctx *ctx, const uint64_t **d_keys, const uint16_t **d_counts, uint64_t *n);

/* ---- hash-range sharding across GPUs (one context per GPU; exchange by the caller with
 * NCCL all-to-all, see metafast_b200/sharded.py) -- no reference analogue --
 * Extract canonical k-mers of a batch and bucket them by owner shard
 * owner(key) = mix(key) mod n_shards.  d_keys_out (device, capacity cap_keys) receives the
 * keys grouped by shard; bucket_counts[n_shards] (host) the size of each group.  */
int  mfkc_extract_bucketed(mfkc_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offsets,
                           uint32_t n_reads, uint64_t n_bases,
                           uint64_t *d_keys_out, uint64_t cap_keys, uint64_t *bucket_counts);
/* Count n keys (device pointer) into this context's table -- the receive side of the exchange. */
int  mfkc_count_keys_device(mfkc_ctx *ctx, const uint64_t *d_keys, uint64_t n);
/* owner shard of a key (host helper, identical to the device function). */
uint32_t mfkc_owner_shard(uint64_t key, uint32_t n_shards);
/* Super-k-mer flavour of the exchange (what bench.py uses for N > 1): the extraction kernel cuts every
 * read into runs of consecutive k-mers that share the owner of their minimizer and ships each run as ONE
 * 16-byte record (up to 16 k-mers: ~2.7 B per k-mer over NVLink instead of 8 B).
 * d_recs_out = n_shards segments of seg_cap records; rec_counts / kmer_counts (host, n_shards entries) =
 * records / k-mer instances per destination.  Returns 1 when a segment overflowed (retry with fewer reads). */
int  mfkc_skm_extract_bucketed(mfkc_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offsets,
                               uint32_t n_reads, uint64_t n_bases, void *d_recs_out, uint64_t seg_cap,
                               uint64_t *rec_counts, uint64_t *kmer_counts);
/* Receive side: file n_recs records (device pointer; n_kmers k-mer instances in total) under their table
 * regions; they are counted at the next drain (mfkc_flush at the latest). */
int  mfkc_skm_count_device(mfkc_ctx *ctx, const void *d_recs, uint64_t n_recs, uint64_t n_kmers);
/* Block until every mfkc_skm_count_device issued so far has finished reading its d_recs buffer (they run on
 * a side stream so that the next extraction overlaps them; call this before overwriting a receive buffer). */
int  mfkc_skm_count_wait(mfkc_ctx *ctx);

/* Peer-memory flavour of the exchange (default of bench.py for N > 1; no reference analogue):
 * every GPU stages its super-k-mer records in its OWN memory, bucketed by (owner shard, coarse bucket of the
 * minimizer hash); the owner's drain kernel then reads its segments straight out of every peer's staging
 * buffer over NVLink (CUDA IPC pointers) while it upserts -- one kernel does transfer + counting, no NCCL
 * call on the data path, no receive buffer, no re-staging.  Call sequence per sample, on every rank:
 *   mfkc_p2p_stage_create  once: (n_shards << log2_buckets) segments of seg_cap 16-byte records
 *   mfkc_p2p_export        -> 128 opaque bytes (two CUDA IPC handles); exchange them between the processes
 *   mfkc_p2p_attach        once per peer (own rank included); mfkc_p2p_attach_ctx for contexts of the same process
 *   [barrier] mfkc_p2p_stage_reset
 *   mfkc_p2p_extract ...   any number of batches (device-resident reads)
 *   mfkc_p2p_counts        -> k-mer instances staged for every owner so far; exchange them (this is also the
 *                             barrier: a rank's counts exist only after its extraction has finished)
 *   mfkc_p2p_drain(n_in)   n_in = k-mer instances staged for this rank on all ranks together
 *   mfkc_flush / mfkc_emit_* as on one GPU.
 * A staging segment that overflows is reported by mfkc_flush (MFKC_E_STATE); nothing is dropped silently.
 * Works for 64-bit and 128-bit keys (k <= 63; 32-byte records for k > 31); up to 16 shards. */
int  mfkc_p2p_stage_create(mfkc_ctx *ctx, uint32_t log2_buckets, uint64_t seg_cap);
/* The same exchange for the bin-local count (MFKC_VARIANT_HASH, k <= 31; the default of bench.py and mfkc_cli --gpus):
 * segments are (owner shard, minimizer bin), bins_per_shard bins per shard sized so that one bin's distinct k-mers fit a
 * shared-memory table, plus an overflow list of ovf_cap records.  mfkc_p2p_drain then only notes the k-mer total; the
 * result calls (mfkc_stats / mfkc_histogram / mfkc_emit_begin) run the counting kernel, which streams this shard's bins
 * out of every peer's staging buffer over NVLink (TMA bulk copies from peer memory) straight into shared memory.  The
 * peers' buffers must stay untouched (no mfkc_p2p_stage_reset) until every rank has fetched its results. */
int  mfkc_p2p_stage_create_bins(mfkc_ctx *ctx, uint32_t bins_per_shard, uint64_t seg_cap, uint64_t ovf_cap);
/* A geometry for mfkc_p2p_stage_create_bins from the expected k-mer instances per rank (host arithmetic only; every rank
 * must pass the same numbers).  instances_per_distinct / slack: 0 = defaults (3.0 / 2.0). */
int  mfkc_p2p_bin_geometry(uint64_t kmers_per_rank, uint32_t n_shards, int k, double instances_per_distinct, double slack,
                           uint32_t *bins_per_shard, uint64_t *seg_cap, uint64_t *ovf_cap);
int  mfkc_p2p_export(mfkc_ctx *ctx, uint8_t handles[128]);
int  mfkc_p2p_attach(mfkc_ctx *ctx, uint32_t rank, const uint8_t handles[128]);
int  mfkc_p2p_attach_ctx(mfkc_ctx *ctx, uint32_t rank, mfkc_ctx *peer);
int  mfkc_p2p_stage_reset(mfkc_ctx *ctx);
int  mfkc_p2p_extract(mfkc_ctx *ctx, const uint8_t *d_bases, const uint64_t *d_offsets, uint32_t n_reads, uint64_t n_bases);
/* host-buffer version of mfkc_p2p_extract: same arguments, copies and overlap as mfkc_submit_reads */
int  mfkc_p2p_submit_reads(mfkc_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads);
int  mfkc_p2p_counts(mfkc_ctx *ctx, uint64_t *kmers_per_owner /* n_shards entries */);
int  mfkc_p2p_drain(mfkc_ctx *ctx, uint64_t n_kmers_in);

/* ---- features-calculator: replaces the BigLong2LongHashMap set-up
 * (src/tools/FeaturesCalculatorMain.java:97-103), IOUtils.calculatePresenceForKmers /
 * ...ForReads (src/io/IOUtils.java:577-597, 806-834) and buildAndPrintVector
 * (src/tools/FeaturesCalculatorMain.java:169-236) -------------------------------------- */
/* keys = all component k-mers, component c = keys[comp_offsets[c]..comp_offsets[c+1]). */
int  mfkc_fc_load_components(mfkc_ctx *ctx, const int64_t *keys, const uint64_t *comp_offsets, uint32_t n_comp);
/* --selected: records of the selected .kmers.bin files (loadKmers with threshold 0,
 * src/io/IOUtils.java:237-258,369-401).  records = NULL switches the filter off; a non-NULL
 * pointer with n_records = 0 is an (empty) active filter, as an empty --selected file is in the
 * reference.  May be called repeatedly to append. */
int  mfkc_fc_set_selected(mfkc_ctx *ctx, const uint8_t *be_records, uint64_t n_records);
int  mfkc_fc_reset_values(mfkc_ctx *ctx);          /* hm.resetValues(), FeaturesCalculatorMain.java:122,151 */
/* 10-byte BE records in any chunking (the reference reads 16 777 200-byte chunks, IOUtils.java:30). */
int  mfkc_fc_add_records(mfkc_ctx *ctx, const uint8_t *be_records, uint64_t n_records);
/* The same for the records a counter context on the same GPU has just selected (after its mfkc_emit_begin): the pairs
 * are taken from the device arrays, no copy to the host and back (kmer-counter-many -> features-calculator in one run). */
int  mfkc_fc_add_emitted(mfkc_ctx *ctx, mfkc_ctx *counter);
/* reads mode (-i): same arguments as mfkc_submit_reads; no minSeqLen on this path. */
int  mfkc_fc_add_reads(mfkc_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t n_reads);
/* per component: vec = sum of values > threshold, found = how many such, cnt = keys considered.
 * breadth = (double)found / cnt is left to the host (Java formatting). */
int  mfkc_fc_features(mfkc_ctx *ctx, int64_t threshold, int64_t *vec, uint64_t *found, uint64_t *cnt);

/* ---- set algebra over .kmers.bin files: the (k-mer -> short) maps behind kmers-filter,
 * unique-kmers-multi and kmers-samples-counter, kept on the device as key-sorted arrays ----------------------------
 * mfkc_kset = one BigLong2ShortHashMap of those tools.  Values are Java shorts with Java's arithmetic: addAndBound
 * saturates at 32767 ([itmo]/utils/NumUtils.java:21-26), put(get + x) wraps, getWithZero reads an absent key and a
 * stored -1 as 0 ([itmo]/structures/map/Long2ShortHashMap.java:160-183).  Output is in ascending key order (the
 * reference writes hash-map iteration order; the order is not part of the format). */
typedef struct mfkc_kset mfkc_kset;
int  mfkc_kset_create(mfkc_ctx *ctx, mfkc_kset **out);
void mfkc_kset_destroy(mfkc_kset *ks);
/* IOUtils.loadKmers (src/io/IOUtils.java:369-401; Kmers2HMWorker.processKmer :249-257): records with freq >
 * freq_threshold are added with addAndBound.  Any chunking and any number of files; mfkc_kset_load_finish after the last. */
int  mfkc_kset_load_records(mfkc_kset *ks, const uint8_t *be_records, uint64_t n_records, int32_t freq_threshold);
int  mfkc_kset_load_finish(mfkc_kset *ks);
int  mfkc_kset_size(mfkc_kset *ks, uint64_t *n);                     /* hm.size() */
int  mfkc_kset_reset_values(mfkc_kset *ks);                          /* hm.resetValues(), src/tools/KmersSamplesCounter.java:92 */
/* The entry loops of the tools, for every (key, v) of src with v > thr:
 *   MFKC_KSET_ADD   dst.put(key, (short)(dst.getWithZero(key) + v))   src/tools/UniqueKmersMultipleSamplesFinder.java:106-108
 *   MFKC_KSET_INC   dst.put(key, (short)(dst.getWithZero(key) + 1))   same :109; src/tools/KmersSamplesCounter.java:102-105
 *   MFKC_KSET_ZERO  if (dst.get(key) > thr) dst.put(key, 0)            src/tools/UniqueKmersMultipleSamplesFinder.java:126-129 */
enum { MFKC_KSET_ADD = 0, MFKC_KSET_INC = 1, MFKC_KSET_ZERO = 2 };
int  mfkc_kset_update(mfkc_kset *dst, const mfkc_kset *src, int op, int32_t thr);
/* IOUtils.filterAndPrintKmers (src/io/IOUtils.java:101-123): entries of hm with value > threshold and
 * filter.getWithZero(key) > filter_threshold; filter = NULL keeps only the first condition (the selection of
 * IOUtils.printKmers, src/io/IOUtils.java:61).  *n_good = how many; fetch them as 10-byte BE records with _next. */
int  mfkc_kset_select_begin(mfkc_kset *hm, const mfkc_kset *filter, int32_t threshold, int32_t filter_threshold, uint64_t *n_good);
int  mfkc_kset_select_next(mfkc_kset *hm, uint8_t *out, size_t cap, size_t *written);
/* hist[v] = number of entries with value v, all entries (the statistics of IOUtils.printKmers, src/io/IOUtils.java:59) */
int  mfkc_kset_histogram(mfkc_kset *ks, uint64_t hist[MFKC_HIST_BINS]);

/* seq-builder on a map: SequencesFinders.thresholdStrategy + AddSequencesShiftingRightTask
 * (src/algo/SequencesFinders.java:13-31, src/algo/AddSequencesShiftingRightTask.java:39-123,
 * src/algo/HashMapOperations.java:13-47): the simple paths of the de Bruijn graph of the k-mers with value >
 * freq_threshold that are at least len_threshold bases long, each once (the orientation whose start k-mer is the
 * smaller one).  Output order = ascending (start k-mer, orientation); the reference's order is thread timing.
 * _begin computes them (*n_sequences, *n_bases = total length); _fetch copies them out and releases them:
 * offsets[n+1] into `bases` ('A','G','C','T'), av/min/max_weight as in structures.Sequence (may be NULL). */
int  mfkc_kset_sequences_begin(mfkc_kset *hm, int32_t freq_threshold, int32_t len_threshold, uint64_t *n_sequences, uint64_t *n_bases);
int  mfkc_kset_sequences_fetch(mfkc_kset *hm, uint64_t *offsets, char *bases, uint32_t *av_weight, uint32_t *min_weight, uint32_t *max_weight);

/* component-cutter's graph half on a map (the main caller's next stage): ComponentsBuilder.splitStrategy
 * (src/algo/ComponentsBuilder.java:24-31,58-84,157-181,198-269; neighbours: src/algo/KmerOperations.java:9-26) over the
 * map that IOUtils.loadReads(sequences, k, minLen) returns (src/tools/ComponentCutterMain.java:81-97).  Connected
 * components of the k-mers with value > 0; fewer than min_component_size k-mers: dropped; up to max_component_size:
 * kept with usedFreqThreshold = the level that found them (1 first); larger: split again among their k-mers of value >=
 * level + 1.  Order = ConnectedComponent.compareTo (src/structures/ConnectedComponent.java:125-136: threshold ascending,
 * weight descending, size descending), ties by ascending smallest k-mer; k-mers ascending inside a component (the
 * reference's BFS / hash-map order is not part of the format).  _begin computes (*n_components, *n_kmers = total
 * members); _fetch copies out and releases: comp_offsets[n_components + 1] into keys, weights[n] =
 * ConnectedComponent.weight, thresholds[n] = usedFreqThreshold (either may be NULL).  Maps of up to 2^32 - 2 entries. */
int  mfkc_kset_components_begin(mfkc_kset *hm, int64_t min_component_size, int64_t max_component_size, uint64_t *n_components, uint64_t *n_kmers);
int  mfkc_kset_components_fetch(mfkc_kset *hm, uint64_t *comp_offsets, int64_t *keys, int64_t *weights, int32_t *thresholds);

/* ---- host side of the path (CPU; no GPU needed): the parser rules of
 * [itmo]/io/ReadersUtils.java:27-102, readers/FastaReader.java:54-108,
 * readers/FastqReader.java:53-114, readers/FastaReaderFromXQSource.java:62-85 and the
 * naming rules of src/tools/KmersCounterForManyFilesMain.java:73-108,
 * src/tools/KmersCounterMain.java:122-137 ---------------------------------------------- */
typedef struct mfkc_reader mfkc_reader;
int  mfkc_reader_open(const char *path, mfkc_reader **out, char *err, size_t err_cap);
/* Fill `bases` (cap_bases bytes) and `offsets` (cap_reads+1 entries) with the next kept reads;
 * *n_reads = 0 at end of file.  A read longer than cap_bases is an MFKC_E_BADARG that loses nothing: the read stays
 * pending, mfkc_reader_pending_bases tells its length, and the next call with a buffer of at least that size returns it
 * (the reference takes FASTA records of any length, e.g. a whole chromosome: FastaReader.java:54-108). */
int  mfkc_reader_next(mfkc_reader *r, uint8_t *bases, size_t cap_bases, uint64_t *offsets,
                      uint32_t cap_reads, uint32_t *n_reads);
/* length of the read the next mfkc_reader_next call starts with, if it is already known (0 otherwise) */
int  mfkc_reader_pending_bases(const mfkc_reader *r, uint64_t *n_bases);
/* counters: [0] = records seen, [1] = records dropped (N / phred 0) */
int  mfkc_reader_counters(const mfkc_reader *r, uint64_t counters[2]);
const char *mfkc_reader_error(const mfkc_reader *r);
const char *mfkc_reader_name(const mfkc_reader *r);   /* NamedSource.name() */
/* the same name from the path alone (file name minus .fasta/.fa/.fn/.fna/.fastq/.fq and .gz): no reader, no threads.
 * MFKC_E_FORMAT for a file whose format ReadersUtils.detectFileFormat would not accept. */
int  mfkc_library_name(const char *path, char *out, size_t cap);
void mfkc_reader_close(mfkc_reader *r);

/* Deterministic merge of per-shard record streams into one .kmers.bin image (multi-GPU runs; no reference analogue,
 * SURVEY.md 8e).  parts[i] = n_records[i] records of record_size bytes (10, or 18 for k > 31), each part in ascending
 * big-endian key order; the shards own disjoint key sets, so the result is the ascending interleave.  `out` holds the sum
 * of all records.  threads <= 0: all host threads.  CPU only. */
int  mfkc_merge_records(const uint8_t *const *parts, const uint64_t *n_records, uint32_t n_parts, uint32_t record_size,
                        uint8_t *out, int threads);

/* .stat.txt text of QuickQuantitativeStatistics.printToFile (header of IOUtils.java:69). */
int  mfkc_write_stat_file(const char *path, const uint64_t hist[MFKC_HIST_BINS]);

/* ---- synthetic Illumina-like reads (bench / tests).  Deterministic
 * in (seed, sample, read index): host and device generators agree byte for byte. ---------- */
typedef struct mfkc_synth_cfg {
    uint32_t struct_size;
    uint32_t n_genomes;        /* default 64 */
    uint64_t seed;             /* default 0x4D464B43 */
    uint64_t total_genome_bp;  /* default 150 000 000 */
    uint32_t read_len;         /* default 150 */
    uint32_t sample;           /* abundance vector id */
    uint32_t err_ppm_first;    /* substitution rate at read start, ppm (default 1000)  */
    uint32_t err_ppm_last;     /* substitution rate at read end, ppm (default 10000)   */
    uint32_t n_read_ppm;       /* reads that get one 'N' (default 1000)                */
    uint32_t poly_tail_ppm;    /* reads with a poly-A / poly-G tail (default 100)      */
    uint64_t reserved[4];
} mfkc_synth_cfg;
void mfkc_synth_defaults(mfkc_synth_cfg *cfg);
/* Host generator: reads [first_read, first_read+n_reads) as fixed-stride ASCII (read_len bytes
 * each, may contain 'N'). */
int  mfkc_synth_reads_host(const mfkc_synth_cfg *cfg, uint64_t first_read, uint64_t n_reads, uint8_t *out);
/* Device generator: only the reads WITHOUT 'N' (what the parser would keep), densely packed
 * into d_bases (capacity n_reads*read_len); *n_kept reads written.  d_offsets (n_reads+1) gets
 * the offsets.  Uses the context's device and stream. */
int  mfkc_synth_reads_device(mfkc_ctx *ctx, const mfkc_synth_cfg *cfg, uint64_t first_read, uint64_t n_reads,
                             uint8_t *d_bases, uint64_t *d_offsets, uint64_t *n_kept);

/* ---- raw device memory helpers so that a non-CUDA host (Java, ctypes) can stage
 * device-resident inputs; thin wrappers over cudaMalloc/cudaFree/cudaMemcpy. ------------- */
int  mfkc_device_alloc(mfkc_ctx *ctx, size_t bytes, void **d_ptr);
int  mfkc_device_free(mfkc_ctx *ctx, void *d_ptr);
int  mfkc_memcpy_h2d(mfkc_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int  mfkc_memcpy_d2h(mfkc_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int  mfkc_device_sync(mfkc_ctx *ctx);
/* CUDA-event timing on the context's streams: marks bracket work submitted in between. */
int  mfkc_timer_start(mfkc_ctx *ctx);
int  mfkc_timer_stop_ms(mfkc_ctx *ctx, float *ms);           /* synchronises */
/* Kernel time/launch accounting since the last reset_profile: name-indexed (see
 * mfkc_profile_name).  Each slot holds accumulated CUDA-event milliseconds (only when
 * profiling is enabled -- it serialises the pipeline) and launch counts (always). */
int  mfkc_profile_enable(mfkc_ctx *ctx, int on);
int  mfkc_profile_reset(mfkc_ctx *ctx);
int  mfkc_profile_get(mfkc_ctx *ctx, int slot, double *ms, uint64_t *launches);
const char *mfkc_profile_name(int slot);                      /* NULL past the last slot */

/* Random-sector microbenchmark (GUPS-style) on a table of `bytes` bytes: n_updates accesses to
 * uniformly random 32-byte sectors; returns device milliseconds.  It measures the random-access
 * HBM roofline the hash variant is compared against.
 *   mode 0: red.add only          mode 1: dependent 128-bit load + red.add (= one table upsert)
 *   mode 2: 128-bit load only     mode 3: mode 1 swept window by window (window_bytes, with
 *                                         blocks_per_window CTAs each): the region-blocked pattern
 * mfkc_gups() is mode 1 over the whole table. */
int  mfkc_gups(mfkc_ctx *ctx, uint64_t bytes, uint64_t n_updates, float *ms);
int  mfkc_gups_ex(mfkc_ctx *ctx, uint64_t bytes, uint64_t n_updates, int mode, uint64_t window_bytes,
                  uint32_t blocks_per_window, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* MFKC_H */
