/*
 * ref_cpu.c -- CPU restatement (plain C, pthreads) of MetaFast's k-mer counting
 * hot path.  TEST INFRASTRUCTURE ONLY: nothing under metafast_b200/ links,
 * loads or executes this file.  It is (1) the second, independent oracle that
 * the numpy oracle (oracle/oracle.py) is cross-checked against, and (2) the
 * "port" CPU baseline that bench.py times on the GPU box's host cores
 * (cpu_baseline.kind = "port", and `bench.py --impl reference`).
 *
 * It follows the reference's ALGORITHM, not just its results: P worker threads
 * pulling batches of <= 32768 reads from a mutex-guarded dispatcher that also
 * does the 2-bit packing (the reference parses+packs inside a synchronized
 * method), rolling canonical k-mers, and a lock-striped set of linear-probing
 * open-addressing sub-maps growing x2 at load 0.75.
 *
 * Citations: src/ = /root/reference/src/ ; [itmo]/ =
 * /root/reference/lib/itmo-assembler-src.jar!/ru/ifmo/genetics/ .
 *
 * Parity status: pinned, transitively, by the reference's own fixture
 * test_data/meta_test_matrix.txt (the matrix-builder result on
 * test_data/meta_test_{1,2,3}.fa): with this file doing the parse + count +
 * filtered emit of every sample, the minSeqLen count of component-cutter and the
 * feature sums, the pipeline reproduces the reference's three Bray-Curtis
 * distances to the last bit (tests/test_oracle.py::test_reference_matrix_golden;
 * see oracle/oracle.py's header for what the fixture does not reach).  No JVM is
 * available, so the reference itself cannot be run.
 * fastutil's HashCommon.murmurHash3 (binary-only dependency, version not pinned
 * in the tree) is restated from the public MurmurHash3 finalizers; it affects
 * slot placement only, never the (key,count) set.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

#define ORC_MAX_COUNT 32767              /* [itmo]/utils/NumUtils.java:21-26 */
#define READS_WORK_RANGE_SIZE (1 << 15)  /* src/io/IOUtils.java:29 */
#define FREE_KEY 0LL                     /* [itmo]/structures/set/LongHashSet.java:33 */

/* ---------------------------------------------------------------- hashing */
/* it.unimi.dsi.fastutil.HashCommon.murmurHash3(int) / (long) */
static inline uint32_t fmix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

/* ------------------------------------------------------- striped hash map */
/* [itmo]/structures/map/Long2ShortHashMap.java (MapData, addAndBound,
 * enlargeAndRehash) on top of [itmo]/structures/set/LongHashSet.java */
typedef struct map_data {
    int64_t *keys;
    int16_t *values;
    int capacity, mask, max_fill;
    volatile int size;
    volatile int contains_free;
    int16_t value_for_free;
    struct map_data *retired;            /* Java GC keeps old arrays alive for racing readers */
} map_data;

typedef struct {
    map_data *volatile data;
    pthread_mutex_t lock;                /* ReentrantLock writeLock, LongHashSet.java:63 */
} small_map;

typedef struct orc_map {
    small_map *maps;
    int n_maps, mask;                    /* BigLong2ShortHashMap.java:44-59 */
    uint64_t total_seq, good_seq, total_len, good_len;   /* IOUtils.java:752-753 */
} orc_map;

static map_data *map_data_new(int capacity) {
    map_data *d = (map_data *)calloc(1, sizeof(map_data));
    d->keys = (int64_t *)calloc((size_t)capacity, sizeof(int64_t));   /* filled with FREE */
    d->values = (int16_t *)calloc((size_t)capacity, sizeof(int16_t));
    d->capacity = capacity;
    d->mask = capacity - 1;
    d->max_fill = (int)ceil(capacity * 0.75f);                        /* LongHashSet.java:57 */
    return d;
}

static void map_data_free(map_data *d) {
    while (d) {
        map_data *r = d->retired;
        free(d->keys); free(d->values); free(d);
        d = r;
    }
}

/* LongHashSet.getPositionInt :153-165 -- lock-free optimistic probe */
static inline int get_position(const map_data *d, int64_t key) {
    int pos = (int)(fmix64((uint64_t)key) & (uint64_t)d->mask);
    for (;;) {
        int64_t cur = __atomic_load_n(&d->keys[pos], __ATOMIC_RELAXED);
        if (cur == FREE_KEY || cur == key) return pos;
        if (++pos == d->capacity) pos = 0;
    }
}

/* Long2ShortHashMap.enlargeAndRehash :191-214 (called under the lock) */
static void enlarge_and_rehash(small_map *m) {
    map_data *cur = m->data;
    map_data *nd = map_data_new(2 * cur->capacity);
    for (int i = 0; i < cur->capacity; i++) {
        int64_t key = cur->keys[i];
        if (key != FREE_KEY) {
            int pos = get_position(nd, key);
            nd->keys[pos] = key;
            nd->values[pos] = cur->values[i];
        }
    }
    nd->contains_free = cur->contains_free;
    nd->value_for_free = cur->value_for_free;
    nd->size = cur->size;
    nd->retired = cur;
    __atomic_store_n(&m->data, nd, __ATOMIC_RELEASE);
}

static inline int16_t add_and_bound16(int16_t v, int16_t inc) {   /* NumUtils.java:21-26 */
    if (v > ORC_MAX_COUNT - inc) return ORC_MAX_COUNT;
    return (int16_t)(v + inc);
}

/* Long2ShortHashMap.addAndBound :119-157 */
static void small_add_and_bound(small_map *m, int64_t key, int16_t inc) {
    if (key == FREE_KEY) {
        pthread_mutex_lock(&m->lock);
        map_data *d = m->data;
        d->value_for_free = add_and_bound16(d->value_for_free, inc);
        if (!d->contains_free) { d->contains_free = 1; d->size++; }
        pthread_mutex_unlock(&m->lock);
        return;
    }
    for (;;) {
        map_data *d = __atomic_load_n(&m->data, __ATOMIC_ACQUIRE);
        int pos = get_position(d, key);
        pthread_mutex_lock(&m->lock);
        if (d == m->data && (d->keys[pos] == FREE_KEY || d->keys[pos] == key)) {
            d->values[pos] = add_and_bound16(d->values[pos], inc);
            if (d->keys[pos] == FREE_KEY) {
                __atomic_store_n(&d->keys[pos], key, __ATOMIC_RELAXED);
                d->size++;
                if (d->size >= d->max_fill) enlarge_and_rehash(m);
            }
            pthread_mutex_unlock(&m->lock);
            return;
        }
        pthread_mutex_unlock(&m->lock);
    }
}

orc_map *orc_map_new(int P) {
    /* src/io/IOUtils.java:775-776: new BigLong2ShortHashMap((int)(log P / log 2) + 4, 12) */
    int log_maps = (int)(log((double)P) / log(2.0)) + 4;
    orc_map *hm = (orc_map *)calloc(1, sizeof(orc_map));
    hm->n_maps = 1 << log_maps;
    hm->mask = hm->n_maps - 1;
    hm->maps = (small_map *)calloc((size_t)hm->n_maps, sizeof(small_map));
    for (int i = 0; i < hm->n_maps; i++) {
        hm->maps[i].data = map_data_new(1 << 12);
        pthread_mutex_init(&hm->maps[i].lock, NULL);
    }
    return hm;
}

void orc_map_free(orc_map *hm) {
    if (!hm) return;
    for (int i = 0; i < hm->n_maps; i++) {
        map_data_free(hm->maps[i].data);
        pthread_mutex_destroy(&hm->maps[i].lock);
    }
    free(hm->maps);
    free(hm);
}

/* BigLong2ShortHashMap.addAndBound :68-71 */
static inline void big_add_and_bound(orc_map *hm, int64_t key, int16_t inc) {
    int n = (int)(fmix32((uint32_t)(int32_t)key) & (uint32_t)hm->mask);
    small_add_and_bound(&hm->maps[n], key, inc);
}

uint64_t orc_map_size(const orc_map *hm) {   /* BigLong2ShortHashMap.java:92-98 */
    uint64_t s = 0;
    for (int i = 0; i < hm->n_maps; i++) s += (uint64_t)hm->maps[i].data->size;
    return s;
}

void orc_map_read_stats(const orc_map *hm, uint64_t out[4]) {
    out[0] = hm->total_seq; out[1] = hm->good_seq; out[2] = hm->total_len; out[3] = hm->good_len;
}

/* -------------------------------------------------------- 2-bit packing   */
/* [itmo]/dna/DnaTools.java:46-60 : A0 G1 C2 T3 */
static int8_t CODE[256];
static pthread_once_t code_once = PTHREAD_ONCE_INIT;
static void init_code(void) {
    memset(CODE, -1, sizeof CODE);
    CODE['A'] = CODE['a'] = 0; CODE['G'] = CODE['g'] = 1;
    CODE['C'] = CODE['c'] = 2; CODE['T'] = CODE['t'] = 3;
}

/* [itmo]/dna/NucArray.java:9-45: 16 nucleotides per int, nucleotide i in bits
 * 2*(i&15).. of word i>>4 */
typedef struct { uint32_t *words; uint32_t length; } dna_t;

static inline void nuc_set(uint32_t *w, uint32_t i, uint32_t v) { w[i >> 4] |= v << (2 * (i & 15)); }
static inline uint32_t nuc_at(const uint32_t *w, uint32_t i) { return (w[i >> 4] >> (2 * (i & 15))) & 3u; }

/* ------------------------------------------------ dispatcher + workers    */
typedef struct {
    pthread_mutex_t mon;                 /* synchronized getWorkRange, src/io/ReadsDispatcher.java:34 */
    const uint8_t *bases;
    const uint64_t *offsets;
    uint64_t n_reads, next;
    int bad_char;
} dispatcher;

typedef struct {
    dispatcher *disp;
    orc_map *hm;
    int k, min_len;
    uint64_t total_seq, good_seq, total_len, good_len;
} worker;

/* src/io/ReadsDispatcher.java:34-53: pull <= 32768 reads; the reference's
 * iterator builds `new Dna(s)` (2-bit pack, [itmo]/dna/Dna.java:59-64) here,
 * i.e. inside the monitor. */
static uint64_t get_work_range(dispatcher *d, dna_t *out, uint32_t **arena, size_t *arena_cap) {
    pthread_mutex_lock(&d->mon);
    uint64_t first = d->next;
    uint64_t n = d->n_reads - first;
    if (n > READS_WORK_RANGE_SIZE) n = READS_WORK_RANGE_SIZE;
    size_t words = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t len = d->offsets[first + i + 1] - d->offsets[first + i];
        words += (len + 15) / 16;
    }
    if (words > *arena_cap) {
        free(*arena);
        *arena_cap = words + words / 4 + 16;
        *arena = (uint32_t *)malloc(*arena_cap * sizeof(uint32_t));
    }
    memset(*arena, 0, words * sizeof(uint32_t));
    size_t w = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *s = d->bases + d->offsets[first + i];
        uint32_t len = (uint32_t)(d->offsets[first + i + 1] - d->offsets[first + i]);
        out[i].words = *arena + w;
        out[i].length = len;
        for (uint32_t j = 0; j < len; j++) {
            int8_t c = CODE[s[j]];
            if (c < 0) { d->bad_char = 1; c = 0; }   /* IllegalArgumentException in the reference */
            nuc_set(out[i].words, j, (uint32_t)c);
        }
        w += (len + 15) / 16;
    }
    d->next = first + n;
    pthread_mutex_unlock(&d->mon);
    return n;
}

/* src/io/IOUtils.java:756-769 + [itmo]/dna/kmers/ShortKmer.java:122-149 */
static void *worker_run(void *arg) {
    worker *wk = (worker *)arg;
    dna_t *batch = (dna_t *)malloc(sizeof(dna_t) * READS_WORK_RANGE_SIZE);
    uint32_t *arena = NULL;
    size_t arena_cap = 0;
    const int k = wk->k;
    const uint64_t mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    for (;;) {
        uint64_t n = get_work_range(wk->disp, batch, &arena, &arena_cap);
        if (n == 0) break;
        for (uint64_t r = 0; r < n; r++) {
            const dna_t *dna = &batch[r];
            wk->total_seq++;
            wk->total_len += dna->length;
            if ((int64_t)dna->length >= (int64_t)wk->min_len) {
                if (dna->length >= (uint32_t)k) {
                    /* ShortKmer(DnaView(dna,0,k)) : KmerUtils.toLong + reverseComplement */
                    uint64_t fw = 0, rc = 0;
                    for (int j = 0; j < k; j++) {
                        uint64_t c = nuc_at(dna->words, (uint32_t)j);
                        fw = (fw << 2) | c;
                        rc = (rc >> 2) | ((3ULL - c) << (2 * k - 2));
                    }
                    big_add_and_bound(wk->hm, (int64_t)(fw < rc ? fw : rc), 1);
                    for (uint32_t i = (uint32_t)k; i < dna->length; i++) {
                        uint64_t c = nuc_at(dna->words, i);
                        fw = ((fw << 2) | c) & mask;                      /* ShortKmer.java:69 */
                        rc = (rc >> 2) | ((3ULL - c) << (2 * k - 2));     /* ShortKmer.java:70 */
                        big_add_and_bound(wk->hm, (int64_t)(fw < rc ? fw : rc), 1);
                    }
                }
                wk->good_seq++;
                wk->good_len += dna->length;
            }
        }
    }
    free(arena);
    free(batch);
    return NULL;
}

/* src/io/IOUtils.java:772-803 (loadReads) for one already-parsed file.  Can be
 * called repeatedly on the same map (files of one sample go into one map,
 * IOUtils.java:840).  Returns 0, or -2 if a non-ACGT character was met. */
int orc_count_reads(orc_map *hm, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                    int k, int min_len, int P) {
    pthread_once(&code_once, init_code);
    if (k < 1 || k > 31 || P < 1) return -1;
    dispatcher disp;
    memset(&disp, 0, sizeof disp);
    pthread_mutex_init(&disp.mon, NULL);
    disp.bases = bases; disp.offsets = offsets; disp.n_reads = n_reads;
    worker *wk = (worker *)calloc((size_t)P, sizeof(worker));
    pthread_t *th = (pthread_t *)calloc((size_t)P, sizeof(pthread_t));
    for (int i = 0; i < P; i++) {
        wk[i].disp = &disp; wk[i].hm = hm; wk[i].k = k; wk[i].min_len = min_len;
        pthread_create(&th[i], NULL, worker_run, &wk[i]);
    }
    for (int i = 0; i < P; i++) {
        pthread_join(th[i], NULL);
        hm->total_seq += wk[i].total_seq; hm->good_seq += wk[i].good_seq;
        hm->total_len += wk[i].total_len; hm->good_len += wk[i].good_len;
    }
    free(wk); free(th);
    pthread_mutex_destroy(&disp.mon);
    return disp.bad_char ? -2 : 0;
}

/* -------------------------------------------------------------- emit      */
static inline void put_be64(uint8_t *p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (56 - 8 * i)); }
static inline void put_be16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }

typedef struct { uint64_t key; int16_t val; } kv_t;
static int kv_cmp(const void *a, const void *b) {
    uint64_t x = ((const kv_t *)a)->key, y = ((const kv_t *)b)->key;
    return x < y ? -1 : x > y;
}

/* src/io/IOUtils.java:45-71 (printKmers).  hist[c] += 1 for EVERY entry
 * (QuickQuantitativeStatistics), records only for value > threshold.
 * Iteration order: sub-map 0..M-1, slot 0..cap-1, then the FREE key
 * (BigLong2ShortHashMap.java:216-253, Long2ShortHashMap.java:322-366).
 * If sort_by_key != 0 records are emitted in ascending key order instead.
 * out may be NULL (count only).  Returns the number of records ("good"). */
uint64_t orc_emit(const orc_map *hm, int threshold, uint8_t *out, uint64_t out_cap_records,
                  uint64_t *hist /* [32768] or NULL */, int sort_by_key) {
    uint64_t n = orc_map_size(hm);
    kv_t *all = (kv_t *)malloc(sizeof(kv_t) * (n ? n : 1));
    uint64_t m = 0;
    for (int i = 0; i < hm->n_maps; i++) {
        const map_data *d = hm->maps[i].data;
        for (int p = 0; p < d->capacity; p++)
            if (d->keys[p] != FREE_KEY) { all[m].key = (uint64_t)d->keys[p]; all[m].val = d->values[p]; m++; }
        if (d->contains_free) { all[m].key = 0; all[m].val = d->value_for_free; m++; }
    }
    if (sort_by_key) qsort(all, m, sizeof(kv_t), kv_cmp);
    uint64_t good = 0;
    for (uint64_t i = 0; i < m; i++) {
        if (hist) hist[(uint16_t)all[i].val]++;
        if (all[i].val > threshold) {
            if (out && good < out_cap_records) {
                put_be64(out + 10 * good, all[i].key);
                put_be16(out + 10 * good + 8, (uint16_t)all[i].val);
            }
            good++;
        }
    }
    free(all);
    return good;
}

/* QuickQuantitativeStatistics.printToFile :65-72 with IOUtils.java:69 header */
int orc_write_stat(const uint64_t *hist, const char *path) {
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "# k-mer frequency\tnumber of such k-mers\n");
    for (int c = 0; c < 32768; c++)
        if (hist[c]) fprintf(f, "%d\t%llu\n", c, (unsigned long long)hist[c]);
    fprintf(f, "\n");
    fclose(f);
    return 0;
}

/* -------------------------------------------------------------- parsers   */
typedef struct orc_reads {
    uint8_t *bases;
    uint64_t *offsets;     /* n_reads + 1 */
    uint64_t n_reads, cap_reads, n_bases, cap_bases;
    uint64_t all_reads, skipped;
    char err[256];
} orc_reads;

static void reads_push(orc_reads *r, const uint8_t *s, uint64_t len) {
    if (r->n_bases + len > r->cap_bases) {
        r->cap_bases = (r->n_bases + len) * 2 + 1024;
        r->bases = (uint8_t *)realloc(r->bases, r->cap_bases);
    }
    if (r->n_reads + 2 > r->cap_reads) {
        r->cap_reads = r->cap_reads * 2 + 1024;
        r->offsets = (uint64_t *)realloc(r->offsets, r->cap_reads * sizeof(uint64_t));
    }
    memcpy(r->bases + r->n_bases, s, len);
    r->offsets[r->n_reads] = r->n_bases;
    r->n_bases += len;
    r->n_reads++;
    r->offsets[r->n_reads] = r->n_bases;
}

static int ends_with_ci(const char *s, const char *suf) {
    size_t n = strlen(s), m = strlen(suf);
    return n >= m && strcasecmp(s + n - m, suf) == 0;
}

/* [itmo]/io/ReadersUtils.java:27-54.  1 = fasta, 2 = fastq, 0 = unknown */
static int detect_format(const char *path) {
    char name[4096];
    const char *b = strrchr(path, '/');
    snprintf(name, sizeof name, "%s", b ? b + 1 : path);
    if (ends_with_ci(name, ".gz")) name[strlen(name) - 3] = 0;
    if (ends_with_ci(name, ".fastq") || ends_with_ci(name, ".fq")) return 2;
    if (ends_with_ci(name, ".fasta") || ends_with_ci(name, ".fa") || ends_with_ci(name, ".fn") ||
        ends_with_ci(name, ".fna")) return 1;
    return 0;
}

static uint8_t *slurp(const char *path, size_t *len_out) {
    gzFile f = gzopen(path, "rb");      /* transparent for non-gz input; GZIPInputStream otherwise */
    if (!f) return NULL;
    gzbuffer(f, 1 << 20);
    size_t cap = 1 << 22, len = 0;
    uint8_t *buf = (uint8_t *)malloc(cap);
    for (;;) {
        if (cap - len < (1 << 20)) { cap *= 2; buf = (uint8_t *)realloc(buf, cap); }
        int r = gzread(f, buf + len, (unsigned)(cap - len > (1u << 30) ? (1u << 30) : cap - len));
        if (r <= 0) break;
        len += (size_t)r;
    }
    gzclose(f);
    *len_out = len;
    return buf;
}

/* BufferedReader.readLine: terminators \n, \r\n, \r; returns 0 at EOF */
typedef struct { const uint8_t *p, *end; } line_src;
static int next_line(line_src *s, const uint8_t **line, size_t *len) {
    if (s->p >= s->end) return 0;
    const uint8_t *q = s->p;
    while (q < s->end && *q != '\n' && *q != '\r') q++;
    *line = s->p; *len = (size_t)(q - s->p);
    if (q < s->end) {
        if (*q == '\r' && q + 1 < s->end && q[1] == '\n') q += 2; else q += 1;
    }
    s->p = q;
    return 1;
}

/* [itmo]/io/readers/FastaReader.java:54-108 */
static int parse_fasta(orc_reads *out, const uint8_t *buf, size_t len) {
    pthread_once(&code_once, init_code);
    line_src src = { buf, buf + len };
    uint8_t *sb = NULL; size_t sb_len = 0, sb_cap = 0;
    const uint8_t *line; size_t ll;
    int more = 1;
    while (more) {
        sb_len = 0;
        for (;;) {                                    /* readNextDataLine :82-104 */
            if (!next_line(&src, &line, &ll)) { more = 0; break; }
            if (ll > 0 && (line[0] == '>' || line[0] == ';')) { if (sb_len > 0) break; }
            else {
                if (sb_len + ll > sb_cap) { sb_cap = (sb_len + ll) * 2 + 256; sb = (uint8_t *)realloc(sb, sb_cap); }
                memcpy(sb + sb_len, line, ll); sb_len += ll;
            }
        }
        if (sb_len == 0) continue;
        out->all_reads++;
        int hasN = 0;
        for (size_t i = 0; i < sb_len; i++) if (sb[i] == 'N' || sb[i] == 'n') { hasN = 1; break; }
        if (hasN) { out->skipped++; continue; }       /* FastaReader.java:58-60 */
        for (size_t i = 0; i < sb_len; i++)
            if (CODE[sb[i]] < 0) {
                snprintf(out->err, sizeof out->err, "Incorrect nucleotide char: \"%c\"", sb[i]);
                free(sb); return -2;
            }
        reads_push(out, sb, sb_len);
    }
    free(sb);
    return 0;
}

/* [itmo]/io/readers/FastqReader.java:84-110; returns 1 line, 0 EOF, <0 error */
static int fastq_next_data_line(line_src *src, const uint8_t **line, size_t *ll, char *err) {
    const uint8_t *s; size_t n;
    int ok = next_line(src, &s, &n);
    while (ok && n == 0) ok = next_line(src, &s, &n);
    if (!ok) return 0;
    if (!(s[0] == '@' || s[0] == '+')) { snprintf(err, 256, "Unknown structure of fastq file!"); return -3; }
    if (!next_line(src, line, ll)) { snprintf(err, 256, "Unexpected end of file."); return -4; }
    return 1;
}

/* one pass over the records; fmt_lo = 64 (Illumina) or 33 (Sanger).
 * max_records = 1000 and out = NULL for the quality sniff.  Returns 0, -5 on an
 * illegal quality value, other negatives on structural errors. */
static int fastq_pass(orc_reads *out, const uint8_t *buf, size_t len, int fmt_lo, uint64_t max_records, char *err) {
    pthread_once(&code_once, init_code);
    line_src src = { buf, buf + len };
    uint64_t rec = 0;
    while (rec < max_records) {
        const uint8_t *data, *qual; size_t dl, ql;
        int r = fastq_next_data_line(&src, &data, &dl, err);
        if (r == 0) break;
        if (r < 0) return r;
        r = fastq_next_data_line(&src, &qual, &ql, err);
        if (r == 0) { snprintf(err, 256, "Unexpected end of file."); return -4; }
        if (r < 0) return r;
        if (dl != ql) { snprintf(err, 256, "Bad DnaQ record: length of chars and quality is not the same."); return -6; }
        int good = 1;
        for (size_t i = 0; i < dl; i++) {           /* FastqReader.java:70-79 */
            uint8_t ch = data[i];
            if (ch == 'N' || ch == 'n' || ch == '.') { good = 0; continue; }
            if (CODE[ch] < 0) { snprintf(err, 256, "Incorrect nucleotide char: \"%c\"", ch); return -2; }
            int q = qual[i];
            if (q < fmt_lo || q > 126) { snprintf(err, 256, "Invalid quality code char: \"%c\"", q); return -5; }
            /* DnaQBuilder.java:32-35 / DnaQ.java:140-150: phred kept in 6 bits */
            if ((((q - fmt_lo) & 63)) == 0) good = 0; /* FastaReaderFromXQSource.java:66-69 */
        }
        rec++;
        if (out) {
            out->all_reads++;
            if (good) reads_push(out, data, dl); else out->skipped++;
        }
    }
    return 0;
}

void orc_reads_free(orc_reads *r) { if (r) { free(r->bases); free(r->offsets); free(r); } }

/* ReadersUtils.readDnaLazy :85-102.  Never returns NULL; check err[0]. */
orc_reads *orc_parse_file(const char *path) {
    orc_reads *r = (orc_reads *)calloc(1, sizeof(orc_reads));
    r->offsets = (uint64_t *)calloc(1024, sizeof(uint64_t)); r->cap_reads = 1024;
    int fmt = detect_format(path);
    if (!fmt) { snprintf(r->err, sizeof r->err, "Can't detect file format for file '%s'", path); return r; }
    size_t len; uint8_t *buf = slurp(path, &len);
    if (!buf) { snprintf(r->err, sizeof r->err, "can't open %s", path); return r; }
    if (fmt == 1) parse_fasta(r, buf, len);
    else {
        /* ReadersUtils.determineQualityFormat :63-77 */
        char err[256] = "";
        int lo = 64;
        int s = fastq_pass(NULL, buf, len, 64, 1000, err);
        if (s == -5) lo = 33;
        else if (s < 0) { snprintf(r->err, sizeof r->err, "%s", err); free(buf); return r; }
        err[0] = 0;
        if (fastq_pass(r, buf, len, lo, UINT64_MAX, err) < 0) snprintf(r->err, sizeof r->err, "%s", err);
    }
    free(buf);
    return r;
}
const uint8_t *orc_reads_bases(const orc_reads *r) { return r->bases; }
const uint64_t *orc_reads_offsets(const orc_reads *r) { return r->offsets; }
uint64_t orc_reads_count(const orc_reads *r) { return r->n_reads; }
uint64_t orc_reads_nbases(const orc_reads *r) { return r->n_bases; }
const char *orc_reads_error(const orc_reads *r) { return r->err; }

/* ---------------------------------------------------- features-calculator */
/* A plain (single-threaded) long->long map is enough for the checker: the
 * reference's BigLong2LongHashMap differs from the short map only in value
 * width ([itmo]/structures/map/Long2LongHashMap.java:119-183). */
typedef struct { uint64_t *keys; int64_t *vals; uint8_t *used; uint64_t cap, size; } ll_map;

static void ll_init(ll_map *m, uint64_t n) {
    uint64_t cap = 16; while (cap < 2 * n + 16) cap <<= 1;
    m->cap = cap; m->size = 0;
    m->keys = (uint64_t *)calloc(cap, 8); m->vals = (int64_t *)calloc(cap, 8); m->used = (uint8_t *)calloc(cap, 1);
}
static void ll_free(ll_map *m) { free(m->keys); free(m->vals); free(m->used); }
static uint64_t ll_find(const ll_map *m, uint64_t key) {
    uint64_t p = fmix64(key) & (m->cap - 1);
    while (m->used[p] && m->keys[p] != key) p = (p + 1) & (m->cap - 1);
    return p;
}
static inline int64_t add_and_bound64(int64_t v, int64_t inc) {  /* NumUtils.java:27-32, Java wrap */
    int64_t lim = (int64_t)((uint64_t)INT64_MAX - (uint64_t)inc);
    if (v > lim) return INT64_MAX;
    return (int64_t)((uint64_t)v + (uint64_t)inc);
}

/* src/tools/FeaturesCalculatorMain.java:97-103,136-163,169-236 +
 * src/io/IOUtils.java:577-588.  comp_offsets[n_comp+1] indexes comp_keys.
 * records = .kmers.bin bytes (n_records x 10).  selected_records (optional,
 * may be NULL) = concatenated --selected .kmers.bin files, loaded with
 * loadKmers(threshold 0) (IOUtils.java:237-258).  Outputs vec/found/cnt per
 * component. */
int orc_features_kmers(const int64_t *comp_keys, const uint64_t *comp_offsets, uint32_t n_comp,
                       const uint8_t *records, uint64_t n_records,
                       const uint8_t *selected_records, uint64_t n_selected,
                       int threshold, int64_t *vec, uint64_t *found, uint64_t *cnt) {
    uint64_t nk = comp_offsets[n_comp];
    ll_map acc; ll_init(&acc, nk);
    for (uint64_t i = 0; i < nk; i++) {               /* hm.put(kmer, 0) */
        uint64_t p = ll_find(&acc, (uint64_t)comp_keys[i]);
        if (!acc.used[p]) { acc.used[p] = 1; acc.keys[p] = (uint64_t)comp_keys[i]; acc.vals[p] = 0; acc.size++; }
    }
    for (uint64_t i = 0; i < n_records; i++) {        /* KmersPresenceWorker.processKmer */
        const uint8_t *r = records + 10 * i;
        uint64_t key = 0; for (int j = 0; j < 8; j++) key = (key << 8) | r[j];
        int16_t freq = (int16_t)((r[8] << 8) | r[9]);
        uint64_t p = ll_find(&acc, key);
        if (acc.used[p]) acc.vals[p] = add_and_bound64(acc.vals[p], freq);
    }
    ll_map sel = {0}; int have_sel = selected_records != NULL;
    if (have_sel) {
        ll_init(&sel, n_selected);
        for (uint64_t i = 0; i < n_selected; i++) {   /* Kmers2HMWorker.processKmer, threshold 0 */
            const uint8_t *r = selected_records + 10 * i;
            uint64_t key = 0; for (int j = 0; j < 8; j++) key = (key << 8) | r[j];
            int16_t freq = (int16_t)((r[8] << 8) | r[9]);
            if (freq > 0) {
                uint64_t p = ll_find(&sel, key);
                if (!sel.used[p]) { sel.used[p] = 1; sel.keys[p] = key; sel.vals[p] = 0; }
                sel.vals[p] = add_and_bound16((int16_t)sel.vals[p], freq);
            }
        }
    }
    for (uint32_t c = 0; c < n_comp; c++) {           /* buildAndPrintVector :186-204 */
        int64_t kmers = 0; uint64_t kc = 0, kf = 0;
        for (uint64_t i = comp_offsets[c]; i < comp_offsets[c + 1]; i++) {
            uint64_t key = (uint64_t)comp_keys[i];
            int use = 1;
            if (have_sel) { uint64_t p = ll_find(&sel, key); use = sel.used[p] && sel.vals[p] > 0; }
            if (use) {
                uint64_t p = ll_find(&acc, key);
                int64_t v = acc.used[p] ? acc.vals[p] : 0;
                if (v > threshold) { kmers = (int64_t)((uint64_t)kmers + (uint64_t)v); kf++; }
                kc++;
            }
        }
        vec[c] = kmers; found[c] = kf; cnt[c] = kc;
    }
    if (have_sel) ll_free(&sel);
    ll_free(&acc);
    return 0;
}

/* ------------------------------------------------------------------ CLI   */
/* ref_cpu count -k K -b B -p P -o out.kmers.bin -s out.stat.txt file...
 * (one sample; all files into one map, as KmersCounterMain does) */
#ifdef ORC_MAIN
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char **argv) {
    int k = 31, b = 1, P = 1; const char *out = NULL, *st = NULL;
    int i = 1;
    if (argc < 2 || strcmp(argv[1], "count")) { fprintf(stderr, "usage: ref_cpu count -k K -b B -p P [-o kmers.bin] [-s stat.txt] files...\n"); return 2; }
    for (i = 2; i < argc && argv[i][0] == '-'; i += 2) {
        if (i + 1 >= argc) return 2;
        if (!strcmp(argv[i], "-k")) k = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-b")) b = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-p")) P = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-o")) out = argv[i + 1];
        else if (!strcmp(argv[i], "-s")) st = argv[i + 1];
        else return 2;
    }
    if (k <= 0 || k > 31) { fprintf(stderr, "The size of k-mer must be in 1..31.\n"); return 1; }
    orc_map *hm = orc_map_new(P);
    double t_parse = 0, t_count = 0;
    for (; i < argc; i++) {
        double t0 = now_s();
        orc_reads *r = orc_parse_file(argv[i]);
        if (r->err[0]) { fprintf(stderr, "%s: %s\n", argv[i], r->err); return 1; }
        double t1 = now_s();
        if (orc_count_reads(hm, r->bases, r->offsets, r->n_reads, k, 0, P)) return 1;
        double t2 = now_s();
        t_parse += t1 - t0; t_count += t2 - t1;
        orc_reads_free(r);
    }
    uint64_t size = orc_map_size(hm);
    uint64_t *hist = (uint64_t *)calloc(32768, 8);
    uint64_t good = orc_emit(hm, b, NULL, 0, hist, 0);
    if (out) {
        uint8_t *buf = (uint8_t *)malloc(10 * (good ? good : 1));
        memset(hist, 0, 32768 * 8);
        orc_emit(hm, b, buf, good, hist, 1);
        FILE *f = fopen(out, "wb"); if (!f) return 1;
        fwrite(buf, 10, good, f); fclose(f); free(buf);
    }
    if (st) orc_write_stat(hist, st);
    fprintf(stderr, "%llu k-mers found, %llu (%.1f%%) of them is good (not erroneous); parse %.3fs count %.3fs, P=%d\n",
            (unsigned long long)size, (unsigned long long)good, size ? good * 100.0 / size : 0.0, t_parse, t_count, P);
    free(hist);
    orc_map_free(hm);
    return 0;
}
#endif
