"""GPU parity tests: the CUDA path (through the C ABI) against the oracle, bit-exact.

Oracles: oracle/oracle.py (numpy restatement) and oracle/ref_cpu.c (C restatement with the
reference's threading + striped hash maps).  Golden numbers: tests/golden/config1.json.
"""
import json
import os
import struct

import numpy as np
import pytest

import metafast_b200 as m
from oracle import oracle as orc
from tests import _oracle_c
from tests.conftest import GOLDEN, INPUTS

pytestmark = pytest.mark.gpu

VARIANTS = [m.VARIANT_HASH, m.VARIANT_SORT, m.VARIANT_HASH_DIRECT, m.VARIANT_HASH_TABLE]


def run_counter(reads_batches, k, b, variant, min_len=0, **kw):
    with m.KmerCounter(k, min_seq_len=min_len, variant=variant, **kw) as kc:
        for batch in reads_batches:
            kc.submit_reads(batch)
        kc.flush()
        rec = kc.emit(b)
        hist = kc.histogram()
        st = kc.stats()
    return rec, hist, st


def check_against_oracle(reads, k, b, variant, min_len=0, batches=1, **kw):
    reads = list(reads)
    step = max(1, (len(reads) + batches - 1) // batches)
    parts = [reads[i:i + step] for i in range(0, len(reads), step)] or [[]]
    rec, hist, st = run_counter(parts, k, b, variant, min_len, **kw)
    counts = orc.count_reads(reads, k, min_len)
    assert rec == orc.kmers_bin(counts, b, k)
    assert {int(c): int(hist[c]) for c in np.nonzero(hist)[0]} == orc.histogram(counts)
    assert st["distinct"] == len(counts)
    tot, good, tot_len, good_len = orc.read_stats(reads, min_len)
    assert (st["total_seq"], st["good_seq"], st["total_len"], st["good_len"]) == (tot, good, tot_len, good_len)
    assert st["kmers"] == sum(max(0, len(r) - k + 1) for r in reads if len(r) >= min_len)
    return rec


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("n", [1, 2, 3])
def test_config1_golden(built, variant, n):
    """BASELINE config 1: meta_test_{1,2,3}.fa, k=31, default -b 1."""
    gold = json.load(open(os.path.join(GOLDEN, "config1.json")))["meta_test_%d" % n]
    path = os.path.join(INPUTS, "meta_test_%d.fa" % n)
    reads = m.read_file_reads(path)
    assert len(reads) == gold["reads"]
    rec, hist, st = run_counter([reads], 31, 1, variant)
    assert len(rec) == gold["kmers_bin_bytes"]
    assert orc.sha256_hex(rec) == gold["sha256_sorted_records"]
    assert st["distinct"] == gold["distinct"] and st["kmers"] == gold["kmer_instances"]
    assert {str(int(c)): int(hist[c]) for c in np.nonzero(hist)[0]} == gold["hist"]


@pytest.mark.parametrize("variant", VARIANTS)
def test_tinytest_fastq(built, variant):
    reads = m.read_file_reads(os.path.join(INPUTS, "tinytest_A.fastq"))
    rec, hist, st = run_counter([reads], 5, 0, variant)
    keys = [k for k, c in orc.load_kmers_bin(rec)]
    assert keys == [26, 35, 104, 140, 193, 262, 416, 444, 501, 560]
    assert all(c == 1 for k, c in orc.load_kmers_bin(rec))


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("k", [1, 2, 5, 15, 16, 17, 21, 30, 31])
def test_edge_cases(built, variant, k):
    rng = np.random.default_rng(1234 + k)
    reads = []
    for L in list(range(0, 70)) + [150, 151, 255, 256, 257, 1000]:
        reads.append("".join(rng.choice(list("ACGT"), L)))
    reads += ["A" * 64, "T" * 64, "a" * 40, "G" * 33, "C" * 33]          # poly-A/T = key 0, lower case
    reads += ["ACGT" * 20, "AATT" * 12, "GC" * 31, "acgtACGT" * 9]         # palindromic k-mers (fw == rc)
    reads += ["", "A", "AC"]
    check_against_oracle(reads, k, 0, variant)
    check_against_oracle(reads, k, 1, variant, batches=7)
    check_against_oracle(reads, k, 2, variant, min_len=40, batches=3)


@pytest.mark.parametrize("variant", VARIANTS)
def test_saturation(built, variant):
    """counts saturate at 32767 ([itmo]/utils/NumUtils.java:21-26)."""
    k = 9
    read = "ACGTTGCAAGGCTTAACG"
    reads = [read] * 40000 + ["A" * 30] * 2000 + ["ACGTTGCAAGG"] * 10
    rec = check_against_oracle(reads, k, 1, variant, batches=5)
    cs = [c for _, c in orc.load_kmers_bin(rec)]
    assert max(cs) == 32767 and cs.count(32767) >= 10


@pytest.mark.parametrize("variant", VARIANTS)
def test_empty_inputs(built, variant):
    rec, hist, st = run_counter([[]], 31, 1, variant)
    assert rec == b"" and hist.sum() == 0 and st["distinct"] == 0
    rec, hist, st = run_counter([["ACGT"], ["", ""]], 31, 0, variant)      # all reads shorter than k
    assert rec == b"" and st["total_seq"] == 3 and st["kmers"] == 0


@pytest.mark.parametrize("variant", [m.VARIANT_HASH, m.VARIANT_HASH_DIRECT, m.VARIANT_HASH_TABLE])
def test_table_growth(built, variant):
    """the table starts tiny and must grow (Long2ShortHashMap.enlargeAndRehash analogue)."""
    rng = np.random.default_rng(7)
    reads = ["".join(rng.choice(list("ACGT"), 150)) for _ in range(4000)]
    check_against_oracle(reads, 31, 0, variant, batches=16, table_slots=1024)


def _skewed_reads(seed=11, n=6000):
    rng = np.random.default_rng(seed)
    genome = "".join(rng.choice(list("ACGT"), 20000))
    reads = [genome[int(i):int(i) + 120] for i in rng.integers(0, len(genome) - 120, n)]
    reads += ["A" * 100] * 300                       # one very hot key
    reads += ["".join(rng.choice(list("ACGT"), 150)) for _ in range(3000)]      # singletons
    return reads


@pytest.mark.parametrize("knobs,expect", [
    ({}, {}),                                                                        # planned normally: everything in shared memory
    ({"MFKC_BIN_COUNT": "3"}, {"split_passes": 1}),                                 # 3 bins for ~400 k distinct k-mers: passes split by hash
    ({"MFKC_BIN_COUNT": "1"}, {"heavy_entries": 1}),                                # one bin: 32 parts are not enough -> table
    ({"MFKC_BIN_COUNT": "64", "MFKC_BIN_SLACK": "0.5", "MFKC_BIN_OVF": "1000000"}, {"overflow_recs": 1}),   # segments too small: chunks of the overflow pool
    ({"MFKC_BIN_COUNT": "1", "MFKC_BIN_SLACK": "0.1", "MFKC_BIN_OVF": "1000000"}, {"overflow_recs": 1, "heavy_entries": 1}),   # ... of bins that end up in the table
    ({"MFKC_BIN_COUNT": "64", "MFKC_BIN_SLACK": "0.25", "MFKC_BIN_OVF": "100000"}, {"bin_mode": 0}),  # overflow list half full: back to the table
])
def test_bin_local_paths(built, monkeypatch, knobs, expect):
    """MFKC_VARIANT_HASH's bin-local count (bincount.cuh) with its exactness paths forced: split passes, heavy bins
    through the table, the overflow list, and the fall-back of the whole sample to the region-blocked table."""
    for k_, v_ in knobs.items():
        monkeypatch.setenv(k_, v_)
    reads = _skewed_reads()
    k, b = 25, 1
    step = 500
    counts = orc.count_reads(reads, k, 0)
    want = {t: orc.kmers_bin(counts, t, k) for t in (0, b)}
    with m.KmerCounter(k, variant=m.VARIANT_HASH, expected_kmers=sum(len(r) - k + 1 for r in reads)) as kc:
        for rep in range(2):                           # second sample: planned from the first one's statistics
            if rep:
                kc.reset()
            for i in range(0, len(reads), step):
                kc.submit_reads(reads[i:i + step])
            kc.flush()
            rec = kc.emit(b)
            hist = kc.histogram()
            st = kc.stats()
            bs = kc.bin_stats()
            assert rec == want[b]
            assert {int(c): int(hist[c]) for c in np.nonzero(hist)[0]} == orc.histogram(counts)
            assert st["distinct"] == len(counts)
            assert kc.emit(0) == want[0]                                 # another threshold: the count pass runs again
            if rep == 0:
                if "bin_mode" not in expect:
                    assert bs["bin_mode"] == 1
                for name, least in expect.items():
                    assert (bs[name] >= least) if name != "bin_mode" else (bs[name] == least), (name, bs)


def test_bin_local_outgrows_plan(built):
    """a sample far larger than the hint leaves the bin-local mode mid-way (bins_to_table) and stays exact"""
    cfg = m.synth_cfg(total_genome_bp=300000, n_genomes=4)
    raw = m.synth_reads_host(cfg, 0, 30000)
    keep = ~(raw == ord("N")).any(axis=1)
    bases = np.ascontiguousarray(raw[keep]).reshape(-1)
    n = int(keep.sum())
    offsets = (np.arange(n + 1, dtype=np.uint64) * np.uint64(cfg.read_len))
    want_rec, want_hist, want_distinct, want_stats = _oracle_c.count(bases, offsets, 31, 1, P=4)
    with m.KmerCounter(31, variant=m.VARIANT_HASH, expected_kmers=200000) as kc:
        step = 2000
        for s in range(0, n, step):
            e = min(n, s + step)
            kc.submit(bases, offsets[s:e + 1])
        kc.flush()
        assert kc.emit(1) == want_rec
        assert kc.stats()["distinct"] == want_distinct
        assert kc.bin_stats()["bin_mode"] == 0
        assert (kc.histogram() == want_hist).all()


@pytest.mark.parametrize("region_shift,staging_bytes", [(4, 0), (6, 8 * 5000), (8, 8 * 200000), (10, 8 * 100)])
def test_region_blocking_geometry(built, region_shift, staging_bytes):
    """Region-blocked hash variant with many tiny regions, staging buffers that overflow (keys then
    go straight to the table), repeated drains and table growth in between."""
    rng = np.random.default_rng(11)
    genome = "".join(rng.choice(list("ACGT"), 20000))
    reads = [genome[int(i):int(i) + 120] for i in rng.integers(0, len(genome) - 120, 6000)]
    reads += ["A" * 100] * 300                       # one very hot key: a single region takes the surplus
    check_against_oracle(reads, 25, 1, m.VARIANT_HASH, batches=12, table_slots=4096,
                         region_shift=region_shift, staging_bytes=staging_bytes)


@pytest.mark.parametrize("variant", VARIANTS)
def test_random_vs_c_oracle(built, variant):
    """~1.2 M k-mer instances with repeats, checked against the threaded C restatement."""
    cfg = m.synth_cfg(total_genome_bp=200000, n_genomes=8)
    raw = m.synth_reads_host(cfg, 0, 10000)
    keep = ~(raw == ord("N")).any(axis=1)
    bases = np.ascontiguousarray(raw[keep]).reshape(-1)
    n = int(keep.sum())
    offsets = (np.arange(n + 1, dtype=np.uint64) * np.uint64(cfg.read_len))
    want_rec, want_hist, want_distinct, want_stats = _oracle_c.count(bases, offsets, 31, 2, P=4)
    with m.KmerCounter(31, variant=variant) as kc:
        step = 1500
        for s in range(0, n, step):
            e = min(n, s + step)
            kc.submit(bases, offsets[s:e + 1])          # sub-range of one big buffer (offsets[0] != 0)
        kc.flush()
        rec = kc.emit(2)
        hist = kc.histogram()
        st = kc.stats()
    assert rec == want_rec
    assert (hist == want_hist).all()
    assert st["distinct"] == want_distinct
    assert [st["total_seq"], st["good_seq"], st["total_len"], st["good_len"]] == want_stats


def test_bad_nucleotide_is_an_error(built):
    with m.KmerCounter(5) as kc:
        kc.submit_reads(["ACGTACGTNACGT"])
        with pytest.raises(m.MfkcError) as e:
            kc.flush()
        assert e.value.code == -8


def test_bad_arguments(built):
    for k in (0, -3, 64, 100):
        with pytest.raises(m.MfkcError) as e:
            m.KmerCounter(k)
        assert e.value.code == -1
    for variant in (m.VARIANT_SORT, m.VARIANT_HASH_DIRECT):            # k > 31 exists for the default variant only
        with pytest.raises(m.MfkcError):
            m.KmerCounter(32, variant=variant)


@pytest.mark.parametrize("k", [32, 33, 47, 48, 55, 62, 63])
def test_long_kmers_128bit(built, k):
    """32 <= k <= 63 (128-bit keys, 18-byte records): no reference behaviour exists
    (KmersCounterMain.java:70-73); checked against the oracle's definition of the extension."""
    rng = np.random.default_rng(500 + k)
    genome = "".join(rng.choice(list("ACGT"), 6000))
    reads = [genome[int(i):int(i) + int(L)] for i, L in zip(rng.integers(0, 5800, 1500), rng.integers(20, 200, 1500))]
    reads += ["A" * 150, "T" * 150, "acgt" * 40, "AC" * 90, "G" * 64, "", "ACGT"]
    rc = {"A": "T", "C": "G", "G": "C", "T": "A"}
    reads += ["".join(rc[c] for c in reversed(r.upper())) for r in reads[:200]]     # both strands -> same keys
    check_against_oracle(reads, k, 0, m.VARIANT_HASH)
    check_against_oracle(reads, k, 1, m.VARIANT_HASH, batches=5, table_slots=2048, region_shift=6)
    check_against_oracle(reads, k, 2, m.VARIANT_HASH, min_len=100, batches=3, staging_bytes=32 * 50)


def test_reset_between_samples(built):
    a = ["ACGTACGTACGTAAACCCGGGTTT" * 3]
    b = ["TTTTGGGGCCCCAAAATTGGCCAA" * 3]
    with m.KmerCounter(11) as kc:
        kc.submit_reads(a); kc.flush(); ra = kc.emit(0)
        kc.reset()
        kc.submit_reads(b); kc.flush(); rb = kc.emit(0)
    assert ra == orc.kmers_bin(orc.count_reads(a, 11), 0, 11)
    assert rb == orc.kmers_bin(orc.count_reads(b, 11), 0, 11)


def test_synth_device_matches_host(built):
    cfg = m.synth_cfg(total_genome_bp=300000, n_genomes=5, n_read_ppm=20000, poly_tail_ppm=5000)
    n = 3000
    raw = m.synth_reads_host(cfg, 100, n)
    keep = ~(raw == ord("N")).any(axis=1)
    want = np.ascontiguousarray(raw[keep]).reshape(-1)
    import ctypes as C
    with m.KmerCounter(31) as kc:
        d_b = kc.device_alloc(n * cfg.read_len)
        d_o = kc.device_alloc((n + 1) * 8)
        kept = C.c_uint64()
        kc._ck(kc.lib.mfkc_synth_reads_device(kc.h, C.byref(cfg), 100, n, C.c_void_p(d_b), C.c_void_p(d_o), C.byref(kept)))
        assert kept.value == int(keep.sum()) and 0 < kept.value < n
        got = np.empty(kept.value * cfg.read_len, dtype=np.uint8)
        off = np.empty(kept.value + 1, dtype=np.uint64)
        kc.d2h(got, d_b); kc.d2h(off, d_o)
        assert (got == want).all()
        assert (off == np.arange(kept.value + 1, dtype=np.uint64) * np.uint64(cfg.read_len)).all()
        # and count straight from the device-resident batch
        kc.submit_device(d_b, d_o, kept.value, kept.value * cfg.read_len)
        kc.flush()
        rec = kc.emit(1)
        kc.device_free(d_b); kc.device_free(d_o)
    offsets = np.arange(kept.value + 1, dtype=np.uint64) * np.uint64(cfg.read_len)
    want_rec, _, _, _ = _oracle_c.count(want, offsets, 31, 1, P=2)
    assert rec == want_rec


@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_logical_shards_equal_unsharded(built, n_shards):
    """Hash-range sharding with G logical shards on one GPU: bucket -> 'exchange' -> per-shard
    count -> key-ordered merge must equal the unsharded result."""
    cfg = m.synth_cfg(total_genome_bp=100000, n_genomes=4, n_read_ppm=0)
    n = 4000
    raw = m.synth_reads_host(cfg, 0, n)
    bases = np.ascontiguousarray(raw).reshape(-1)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(cfg.read_len)
    want_rec, want_hist, _, _ = _oracle_c.count(bases, offsets, 31, 1, P=2)
    lib = m.load()
    shards = [m.KmerCounter(31, n_shards=n_shards, shard_id=s) for s in range(n_shards)]
    try:
        src = shards[0]
        d_b = src.device_alloc(bases.nbytes); d_o = src.device_alloc(offsets.nbytes)
        src.h2d(d_b, bases); src.h2d(d_o, offsets)
        cap = bases.size
        d_keys = src.device_alloc(cap * 8)
        counts = src.extract_bucketed(d_b, d_o, n, bases.size, d_keys, cap, n_shards)
        assert sum(counts) == n * (cfg.read_len - 30)
        keys = np.empty(sum(counts), dtype=np.uint64)
        src.d2h(keys, d_keys)
        pos = 0
        merged = []
        hist = np.zeros(m.HIST_BINS, dtype=np.uint64)
        for s in range(n_shards):
            part = keys[pos:pos + counts[s]]; pos += counts[s]
            assert all(lib.mfkc_owner_shard(int(x), n_shards) == s for x in part[:200])
            d_part = shards[s].device_alloc(max(part.nbytes, 8))
            if part.size:
                shards[s].h2d(d_part, part)
                shards[s].count_keys_device(d_part, part.size)
            shards[s].flush()
            merged.append(shards[s].emit(1))
            hist += shards[s].histogram()
            shards[s].device_free(d_part)
        # shards hold disjoint key sets: the deterministic merge is a k-way merge by key
        recs = sorted(r for part in merged for r in (part[i:i + 10] for i in range(0, len(part), 10)))
        if b"".join(recs) != want_rec:                  # say what differs (key, got, want, owner shard)
            got_d = {}
            for sh, part in enumerate(merged):
                for key, c in orc.load_kmers_bin(part):
                    got_d.setdefault(key, []).append((c, sh))
            want_d = dict(orc.load_kmers_bin(want_rec))
            diff = [(hex(key), got_d.get(key), want_d.get(key), lib.mfkc_owner_shard(key, n_shards))
                    for key in sorted(set(got_d) | set(want_d)) if [c for c, _ in got_d.get(key, [])] != ([want_d[key]] if key in want_d else [])]
            raise AssertionError("records differ (%d keys): %r; part sizes %r, counts %r" % (len(diff), diff[:8], [len(p_) // 10 for p_ in merged], counts))
        assert (hist == want_hist).all()
        src.device_free(d_b); src.device_free(d_o); src.device_free(d_keys)
    finally:
        for s in shards:
            s.close()


@pytest.mark.parametrize("n_shards", [2, 5])
def test_logical_shards_superkmer_exchange(built, n_shards):
    """The super-k-mer flavour of the exchange with G logical shards on one GPU: records bucketed by the
    owner of their minimizer -> 'exchange' -> filed under table regions -> drained; plus the overflow
    signal of a send segment that is too small."""
    cfg = m.synth_cfg(total_genome_bp=100000, n_genomes=4, n_read_ppm=0)
    n = 4000
    raw = m.synth_reads_host(cfg, 0, n)
    bases = np.ascontiguousarray(raw).reshape(-1)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(cfg.read_len)
    want_rec, want_hist, _, want_stats = _oracle_c.count(bases, offsets, 31, 1, P=2)
    shards = [m.KmerCounter(31, n_shards=n_shards, shard_id=s, table_slots=1 << 16, region_shift=9) for s in range(n_shards)]
    try:
        src = shards[0]
        d_b = src.device_alloc(bases.nbytes); d_o = src.device_alloc(offsets.nbytes)
        src.h2d(d_b, bases); src.h2d(d_o, offsets)
        seg_cap = n * 120
        d_recs = src.device_alloc(n_shards * seg_cap * 16)
        over, _, _ = src.skm_extract_bucketed(d_b, d_o, n, bases.size, d_recs, 100, n_shards)
        assert over                                                   # 100 records per segment cannot hold 480 k k-mers
        over, rc, kmc = src.skm_extract_bucketed(d_b, d_o, n, bases.size, d_recs, seg_cap, n_shards)
        assert not over and sum(kmc) == n * (cfg.read_len - 30) and sum(rc) < sum(kmc) / 3
        st = src.stats()
        assert [st["total_seq"], st["good_seq"], st["total_len"], st["good_len"]] == want_stats   # counted once, not twice
        merged, hist = [], np.zeros(m.HIST_BINS, dtype=np.uint64)
        for s in range(n_shards):
            part = np.empty(rc[s] * 2, dtype=np.uint64)
            src.d2h(part, d_recs + s * seg_cap * 16)
            d_part = shards[s].device_alloc(max(part.nbytes, 16))
            shards[s].h2d(d_part, part)
            shards[s].skm_count_device(d_part, rc[s], kmc[s])
            shards[s].flush()
            merged.append(shards[s].emit(1))
            hist += shards[s].histogram()
            shards[s].device_free(d_part)
        from metafast_b200.sharded import merge_sorted_records
        assert merge_sorted_records(merged) == want_rec
        assert (hist == want_hist).all()
        src.device_free(d_b); src.device_free(d_o); src.device_free(d_recs)
    finally:
        for s in shards:
            s.close()


@pytest.mark.parametrize("n_shards,log2_buckets", [(1, 3), (2, 0), (4, 5), (8, 2), (11, 4)])
def test_logical_shards_peer_memory_exchange(built, n_shards, log2_buckets):
    """Peer-memory flavour of the exchange (mfkc_p2p_*) with G contexts on one GPU: every shard extracts ITS
    slice of the reads into its own staging buffer, every owner drains its segments out of all G staging
    buffers (same kernels as across GPUs; the pointers are attached directly instead of through CUDA IPC).
    Merged records, summed histogram and summed read statistics must equal the unsharded result."""
    cfg = m.synth_cfg(total_genome_bp=100000, n_genomes=4, n_read_ppm=0)
    n = 6000
    raw = m.synth_reads_host(cfg, 0, n)
    bases = np.ascontiguousarray(raw).reshape(-1)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(cfg.read_len)
    want_rec, want_hist, _, want_stats = _oracle_c.count(bases, offsets, 31, 1, P=2)
    from metafast_b200.sharded import merge_sorted_records
    shards = [m.KmerCounter(31, n_shards=n_shards, shard_id=s, table_slots=1 << 15) for s in range(n_shards)]   # small tables: growth on the way
    try:
        seg_cap = 2 * n * 120 // 5 // (n_shards << log2_buckets) + 64
        for s in shards:
            s.p2p_stage_create(log2_buckets, seg_cap)
        for s in shards:
            for r, peer in enumerate(shards):
                s.p2p_attach_ctx(r, peer)
        bounds = np.linspace(0, n, n_shards + 1).astype(int) // 8 * 8
        bounds[-1] = n
        for rep in range(2):                                   # second pass: reset + reuse of the staging buffers
            for s in shards:
                s.reset()
            for s in shards:
                s.p2p_stage_reset()
            bufs = []
            for r, s in enumerate(shards):
                lo, hi = int(bounds[r]), int(bounds[r + 1])
                part_b = bases[lo * cfg.read_len: hi * cfg.read_len]
                part_o = offsets[lo: hi + 1]
                d_b = s.device_alloc(max(part_b.nbytes, 16)); d_o = s.device_alloc(part_o.nbytes)
                if hi > lo and rep == 1:                        # host buffers through the copy pipeline
                    s.p2p_submit(bases, offsets[lo: (lo + hi) // 2 + 1])
                    s.p2p_submit(bases, offsets[(lo + hi) // 2: hi + 1])
                elif hi > lo:
                    s.h2d(d_b, part_b); s.h2d(d_o, part_o)
                    half = (hi - lo) // 2 // 8 * 8              # two batches per shard (device batches start 16-byte aligned)
                    s.p2p_extract(d_b, d_o, half, half * cfg.read_len)
                    s.p2p_extract(d_b + half * cfg.read_len, d_o + half * 8, hi - lo - half, (hi - lo - half) * cfg.read_len)
                bufs.append((s, d_b, d_o))
            counts = [s.p2p_counts(n_shards) for s in shards]
            assert sum(sum(c) for c in counts) == n * (cfg.read_len - 30)
            merged, hist, stats = [], np.zeros(m.HIST_BINS, dtype=np.uint64), np.zeros(4, dtype=np.int64)
            for r, s in enumerate(shards):
                s.p2p_drain(sum(c[r] for c in counts))
                s.flush()
            for s in shards:
                merged.append(s.emit(1))
                hist += s.histogram()
                st = s.stats()
                stats += np.array([st["total_seq"], st["good_seq"], st["total_len"], st["good_len"]])
            assert merge_sorted_records(merged) == want_rec
            assert (hist == want_hist).all()
            assert stats.tolist() == want_stats
            for s, d_b, d_o in bufs:
                s.device_free(d_b); s.device_free(d_o)
    finally:
        for s in shards:
            s.close()


@pytest.mark.parametrize("n_shards,bins,slack", [(1, 40, 2.0), (2, 64, 2.0), (4, 3, 2.0), (8, 16, 0.5), (11, 1, 2.0)])
def test_logical_shards_bin_local_exchange(built, n_shards, bins, slack):
    """The bin-local count over peer memory (mfkc_p2p_stage_create_bins) with G contexts on one GPU: every shard stages
    ITS slice of the reads per (owner, bin); every owner counts its bins in shared memory straight out of all G staging
    buffers.  Few bins force split passes and heavy bins, slack < 1 the overflow lists.  Merged records, summed
    histograms, distinct counts and read statistics must equal the unsharded result."""
    from metafast_b200.sharded import merge_sorted_records
    cfg = m.synth_cfg(total_genome_bp=100000, n_genomes=4, n_read_ppm=0)
    n = 6000
    raw = m.synth_reads_host(cfg, 0, n)
    bases = np.ascontiguousarray(raw).reshape(-1)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(cfg.read_len)
    want_rec, want_hist, want_distinct, want_stats = _oracle_c.count(bases, offsets, 31, 1, P=2)
    shards = [m.KmerCounter(31, n_shards=n_shards, shard_id=s) for s in range(n_shards)]
    try:
        recs_est = n * 120 * 0.19
        seg_cap = int(recs_est / (n_shards * n_shards * bins) * slack) + 8
        for s in shards:
            s.p2p_stage_create_bins(bins, seg_cap, 1 << 18)
        for s in shards:
            for r, peer in enumerate(shards):
                s.p2p_attach_ctx(r, peer)
        bounds = np.linspace(0, n, n_shards + 1).astype(int) // 8 * 8
        bounds[-1] = n
        for rep in range(2):
            for s in shards:
                s.reset()
            for s in shards:
                s.p2p_stage_reset()
            for r, s in enumerate(shards):
                lo, hi = int(bounds[r]), int(bounds[r + 1])
                if hi > lo:
                    s.p2p_submit(bases, offsets[lo: (lo + hi) // 2 + 1])
                    s.p2p_submit(bases, offsets[(lo + hi) // 2: hi + 1])
            counts = [s.p2p_counts(n_shards) for s in shards]
            assert sum(sum(c) for c in counts) == n * (cfg.read_len - 30)
            for r, s in enumerate(shards):
                s.p2p_drain(sum(c[r] for c in counts))
                s.flush()
            merged, hist, stats, distinct = [], np.zeros(m.HIST_BINS, dtype=np.uint64), np.zeros(4, dtype=np.int64), 0
            seen = {"heavy_entries": 0, "split_passes": 0, "overflow_recs": 0}
            for s in shards:
                merged.append(s.emit(1))
                hist += s.histogram()
                st = s.stats()
                distinct += st["distinct"]
                stats += np.array([st["total_seq"], st["good_seq"], st["total_len"], st["good_len"]])
                bs = s.bin_stats()
                assert bs["bin_mode"] == 1
                for key in seen:
                    seen[key] += bs[key]
            assert merge_sorted_records(merged) == want_rec
            assert (hist == want_hist).all()
            assert distinct == want_distinct
            assert stats.tolist() == want_stats
            if bins <= 3:
                assert seen["split_passes"] + seen["heavy_entries"] > 0, seen
            if slack < 1:
                assert seen["overflow_recs"] > 0, seen
    finally:
        for s in shards:
            s.close()


@pytest.mark.parametrize("k,n_shards", [(55, 3), (33, 8)])
def test_peer_memory_exchange_long_kmers(built, k, n_shards):
    """BASELINE config 5 layout in small: 128-bit keys through the peer-memory exchange (18-byte records merged by key)."""
    rng = np.random.default_rng(k)
    genome = "".join(rng.choice(list("ACGT"), 20000))
    L = 144                                                    # 16-byte aligned device batches
    reads = [genome[int(i):int(i) + L] for i in rng.integers(0, 20000 - L, 1600)]
    reads[5] = "A" * L; reads[6] = "T" * L
    want = orc.kmers_bin(orc.count_reads(reads, k), 1, k)
    from metafast_b200.sharded import merge_sorted_records
    bases = np.frombuffer("".join(reads).encode(), dtype=np.uint8).copy()
    offsets = np.arange(len(reads) + 1, dtype=np.uint64) * np.uint64(L)
    shards = [m.KmerCounter(k, n_shards=n_shards, shard_id=s, table_slots=1 << 14) for s in range(n_shards)]
    try:
        for s in shards:
            s.p2p_stage_create(3, 4 * len(reads) * (L - k + 1) // 5 // (n_shards << 3) + 64)
        for s in shards:
            for r, peer in enumerate(shards):
                s.p2p_attach_ctx(r, peer)
            s.p2p_stage_reset()
        per = len(reads) // n_shards // 8 * 8
        for r, s in enumerate(shards):
            lo, hi = r * per, (len(reads) if r == n_shards - 1 else (r + 1) * per)
            s.p2p_submit(bases, offsets[lo: hi + 1])
        counts = [s.p2p_counts(n_shards) for s in shards]
        assert sum(sum(c) for c in counts) == len(reads) * (L - k + 1)
        merged = []
        for r, s in enumerate(shards):
            s.p2p_drain(sum(c[r] for c in counts))
            s.flush()
            merged.append(s.emit(1))
        assert merge_sorted_records(merged, 18) == want
    finally:
        for s in shards:
            s.close()


def test_peer_memory_exchange_reports_overflow(built):
    """a staging segment that is too small must surface as an error at flush, never as a silently short count"""
    cfg = m.synth_cfg(total_genome_bp=100000, n_genomes=4, n_read_ppm=0)
    n = 2000
    bases = np.ascontiguousarray(m.synth_reads_host(cfg, 0, n)).reshape(-1)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(cfg.read_len)
    with m.KmerCounter(31, n_shards=2, shard_id=0) as a, m.KmerCounter(31, n_shards=2, shard_id=1) as b:
        for s in (a, b):
            s.p2p_stage_create(1, 50)
        for s in (a, b):
            s.p2p_attach_ctx(0, a); s.p2p_attach_ctx(1, b)
        d_b = a.device_alloc(bases.nbytes); d_o = a.device_alloc(offsets.nbytes)
        a.h2d(d_b, bases); a.h2d(d_o, offsets)
        a.p2p_extract(d_b, d_o, n, bases.size)
        with pytest.raises(m.MfkcError):
            a.flush()
        a.device_free(d_b); a.device_free(d_o)


# ---------------------------------------------------------------- features-calculator
def _components_from(counts, rng, n_comp=40):
    keys = sorted(counts)
    comps = []
    for _ in range(n_comp):
        size = int(rng.integers(0, 60))
        comp = [keys[int(i)] for i in rng.integers(0, len(keys), size)]
        comp += [int(x) for x in rng.integers(0, 1 << 62, int(rng.integers(0, 5)))]   # keys absent from the sample
        comps.append(comp)
    comps.append([])                        # empty component: breadth NaN
    comps.append([keys[0]] * 5)             # duplicated key
    return comps


def test_features_from_kmers_files(built):
    rng = np.random.default_rng(99)
    path = os.path.join(INPUTS, "meta_test_3.fa")
    reads = orc.parse_reads(path)
    counts = orc.count_reads(reads, 31)
    records = orc.kmers_bin(counts, 1, 31)
    comps = _components_from(counts, rng)
    sel_counts = {k: c for k, c in counts.items() if k % 3 == 0}
    selected = orc.kmers_bin(sel_counts, 0, 31)
    with m.FeaturesCalculator(31) as fc:
        fc.load_components(comps)
        for thr, sel in ((0, None), (3, None), (0, selected), (2, selected), (0, b"")):
            fc.set_selected(sel)
            fc.reset_values()
            fc.add_records(records, chunk=10 * 1000)
            vec, found, cnt = fc.features(thr)
            acc = orc.presence_for_kmers([k for c in comps for k in c], orc.load_kmers_bin(records))
            seld = None if sel is None else orc.load_kmers([sel], 0)
            wv, wb, wf, wc = orc.features([(0, c) for c in comps], acc, thr, seld)
            assert list(vec) == wv and list(found) == wf and list(cnt) == wc
            cv, cf, cc = _oracle_c.features_kmers(comps, records, sel, thr)
            assert list(cv) == wv and list(cf) == wf and list(cc) == wc
            fc.set_selected(None)
        # a file loaded twice adds up (no reset in between); negative 'freq' follows Java's addAndBound
        fc.reset_values()
        fc.add_records(records); fc.add_records(records)
        vec2, _, _ = fc.features(0)
        acc = orc.presence_for_kmers([k for c in comps for k in c], orc.load_kmers_bin(records) * 2)
        assert list(vec2) == orc.features([(0, c) for c in comps], acc, 0)[0]


@pytest.mark.parametrize("variant", [m.VARIANT_HASH, m.VARIANT_SORT])
def test_features_from_the_counters_device_arrays(built, variant):
    """kmer-counter -> features-calculator without the .kmers.bin round trip (mfkc_fc_add_emitted): the same vectors as
    from the records file."""
    rng = np.random.default_rng(123)
    reads = orc.parse_reads(os.path.join(INPUTS, "meta_test_1.fa"))
    counts = orc.count_reads(reads, 31)
    comps = _components_from(counts, rng)
    with m.KmerCounter(31, variant=variant) as kc, m.FeaturesCalculator(31) as fc:
        fc.load_components(comps)
        kc.submit_reads(reads)
        kc.flush()
        for b, thr in ((1, 0), (0, 2)):
            n_good = kc.emit_begin(b)
            records = orc.kmers_bin(counts, b, 31)
            assert n_good == len(records) // 10
            fc.reset_values()
            fc.add_emitted(kc)
            vec, found, cnt = fc.features(thr)
            acc = orc.presence_for_kmers([k for c in comps for k in c], orc.load_kmers_bin(records))
            wv, wb, wf, wc = orc.features([(0, c) for c in comps], acc, thr)
            assert list(vec) == wv and list(found) == wf and list(cnt) == wc


def test_features_from_reads(built):
    rng = np.random.default_rng(5)
    reads = orc.parse_reads(os.path.join(INPUTS, "meta_test_2.fa"))
    counts = orc.count_reads(reads, 21)
    comps = _components_from(counts, rng, 25)
    with m.FeaturesCalculator(21) as fc:
        fc.load_components(comps)
        fc.reset_values()
        fc.add_reads(reads[:300]); fc.add_reads(reads[300:])
        vec, found, cnt = fc.features(0)
    acc = orc.presence_for_reads([k for c in comps for k in c], reads, 21)
    wv, wb, wf, wc = orc.features([(0, c) for c in comps], acc, 0)
    assert list(vec) == wv and list(found) == wf and list(cnt) == wc


# ---------------------------------------------------------------- set algebra over .kmers.bin files
def _random_kmers_files(rng, n_files, universe, per_file, hot):
    """unsorted record files over a shared key universe: duplicates inside a file (addAndBound), values around the
    threshold, and a few keys whose values sit at the short limits (saturation at 32767, (short) wrap, stored -1)"""
    keys = rng.integers(0, 1 << 62, universe, dtype=np.uint64)
    files = []
    for f in range(n_files):
        idx = rng.integers(0, universe, per_file)
        vals = rng.choice([1, 2, 3, 5, 40, 1000, 20000, 32767], per_file)
        recs = [(int(keys[i]), int(v)) for i, v in zip(idx, vals)]
        recs += [(int(keys[j]), 32767) for j in range(hot)]                     # same hot keys in every file
        if f == 2:
            recs += [(int(keys[j]), 1) for j in range(hot)]                     # 32767 + 32767 + ... : exercises -1 and the wrap
        rng.shuffle(recs)
        files.append(b"".join(struct.pack(">Qh", k, v) for k, v in recs))
    return files


@pytest.mark.parametrize("b", [0, 1, 2])
def test_kmer_set_tools(built, b):
    import struct as _s  # noqa: F401
    rng = np.random.default_rng(20 + b)
    inputs = _random_kmers_files(rng, 4, 3000, 2500, 6)
    filters = _random_kmers_files(np.random.default_rng(77), 2, 3000, 1500, 3)
    filters = [f[: len(f) // 10 * 10] for f in filters]
    # filter files share part of the key universe with the inputs
    shared = inputs[0][:4000] + filters[0]
    with m.KmerCounter(31) as ctx:
        # loadKmers
        for thr in (b, 5):
            want = orc.load_kmers(inputs[:2], thr)
            with m.KmerSet.load(ctx, inputs[:2], thr, chunk=7000) as ks:
                assert ks.size() == len(want)
                assert ks.select(None, -1) == orc._records_of(want.items())
        # kmers-filter
        want = orc.kmers_filter(inputs[:2], [shared, filters[1]], b, 0)
        got = m.kmers_filter(ctx, inputs[:2], [shared, filters[1]], b, 0)
        assert got == want and any(len(r) for _, r in want)
        assert m.kmers_filter(ctx, inputs[:1], [shared], b, 3) == orc.kmers_filter(inputs[:1], [shared], b, 3)
        # unique-kmers-multi (incl. the (short) wrap of the value sum and getWithZero's -1 rule)
        want = orc.unique_kmers_multi(inputs, [shared], b, 1, 4)
        got = m.unique_kmers_multi(ctx, inputs, [shared], b, 1, 4)
        assert got[0] == want[0]
        for i in range(1, 5):
            assert got[1][i] == want[1][i], i
        assert len(want[1][1]) > len(want[1][3]) > 0
        # kmers-samples-counter
        n, recs, stat = orc.kmers_samples_counter(inputs, b)
        gn, grecs, ghist = m.kmers_samples_counter(ctx, inputs, b)
        assert gn == n and grecs == recs
        lines = ["# k-mer frequency\tnumber of such k-mers"] + ["%d\t%d" % (c, int(ghist[c])) for c in np.nonzero(ghist)[0]]
        assert "\n".join(lines) + "\n\n" == stat


def test_kmer_set_edge_cases(built):
    with m.KmerCounter(31) as ctx:
        with m.KmerSet.load(ctx, [b""], 1) as empty, m.KmerSet.load(ctx, [struct.pack(">Qh", 7, 5) + struct.pack(">Qh", 0, 9)], 1) as two:
            assert empty.size() == 0 and empty.select(None, 0) == b""
            assert two.size() == 2 and two.select(empty, 1, 0) == b"" and two.select(empty, 1, -1) == two.select(None, 1)
            two.update(empty, m.KmerSet.ADD, 1)
            assert two.size() == 2
            empty.update(two, m.KmerSet.INC, 1)
            assert empty.select(None, 0) == struct.pack(">Qh", 0, 1) + struct.pack(">Qh", 7, 1)     # key 0 (poly-A) is a legal key
            two.update(empty, m.KmerSet.ZERO, 0)
            assert two.select(None, -1) == struct.pack(">Qh", 0, 0) + struct.pack(">Qh", 7, 0)


@pytest.mark.parametrize("k,thr,min_len", [(21, 1, 30), (15, 0, 15), (31, 2, 100), (5, 0, 5)])
def test_seq_builder(built, k, thr, min_len):
    """simple paths of the de Bruijn graph (src/algo/AddSequencesShiftingRightTask.java) against the
    oracle: branching genome (repeats), both strands, palindromic k-mers, isolated k-mers, erroneous low-count k-mers"""
    rng = np.random.default_rng(100 + k)
    unit = "".join(rng.choice(list("ACGT"), 400))
    genome = unit + "".join(rng.choice(list("ACGT"), 1500)) + unit[:250] + "".join(rng.choice(list("ACGT"), 800)) + "ACGT" * 12 + "AT" * 15
    rc = {"A": "T", "C": "G", "G": "C", "T": "A"}
    reads = [genome[int(i):int(i) + 90] for i in rng.integers(0, len(genome) - 90, 900)]
    reads += ["".join(rc[c] for c in reversed(r)) for r in reads[:300]]
    for j in range(0, 60, 7):                                      # sequencing errors: low-count tips and bubbles
        r = list(reads[j]); r[45] = rc[r[45]]; reads.append("".join(r))
    counts = orc.count_reads(reads, k)
    data = orc.kmers_bin(counts, 0, k)
    want = orc.seq_builder(orc.load_kmers([data], thr, k), k, thr, min_len)
    with m.KmerCounter(k) as ctx, m.KmerSet.load(ctx, [data], thr) as ks:
        got = ks.sequences(thr, min_len)
        assert got == want
        assert ks.sequences(thr, 10 ** 6) == []                   # nothing is that long
    assert len(want) >= 1


def test_reference_matrix_golden(built):
    """The reference's only golden output for this path, test_data/meta_test_matrix.txt (the matrix-builder result on
    meta_test_{1,2,3}.fa, k=31, -b 1, -l 100; copied verbatim to tests/golden/), reproduced with the device doing every
    in-scope stage through the C ABI: counting + filtered emit, seq-builder, component-cutter's minSeqLen counting call,
    features.  The component split and the Bray-Curtis formula (out of scope) are the checker's.
    Exact doubles."""
    gold = orc.load_matrix_txt(open(os.path.join(GOLDEN, "meta_test_matrix.txt")).read())
    names = ["meta_test_%d" % n for n in (1, 2, 3)]
    recs, seq_reads = {}, []
    for name in names:
        reads = m.read_file_reads(os.path.join(INPUTS, name + ".fa"))
        recs[name], _, _ = run_counter([reads], 31, 1, m.VARIANT_HASH)
        with m.KmerCounter(31) as ctx, m.KmerSet.load(ctx, [recs[name]], 1) as ks:
            seq_reads += [s[0] for s in ks.sequences(1, 100)]
    assert len(seq_reads) == 15 + 29 + 25
    seq_rec, _, st = run_counter([seq_reads], 31, 0, m.VARIANT_HASH, min_len=100)
    seq_hm = dict(orc.load_kmers_bin(seq_rec))
    assert st["distinct"] == len(seq_hm) == 17061
    comps3 = orc.component_cutter(seq_hm, 31, 1000, 10000)
    vecs = {}
    with m.FeaturesCalculator(31) as fc:
        fc.load_components([keys for _w, keys, _thr in comps3])
        for name in names:
            fc.reset_values()
            fc.add_records(recs[name])
            vecs[name] = [int(x) for x in fc.features(0)[0]]
    assert vecs["meta_test_1"] == [41935, 38354, 20375, 14211]
    for a in names:
        for b in names:
            if a != b:
                assert orc.bray_curtis(vecs[a], vecs[b]) == gold[(a, b)], (a, b)


@pytest.mark.parametrize("k,b1,b2", [(31, 50, 800), (21, 1, 100), (11, 30, 2000), (5, 1, 20), (31, 1000, 10000)])
def test_component_cutter(built, k, b1, b2):
    """component-cutter's graph half (mfkc_kset_components_*, csrc/components.cuh) against the restatement of
    ComponentsBuilder (src/algo/ComponentsBuilder.java:24-31): size window, re-split of big components level by level
    (up to 264 levels at k = 5), weights, thresholds, order"""
    reads = []
    for n in ((1, 2, 3) if b1 == 1000 else (2,)):
        reads += orc.parse_reads(os.path.join(INPUTS, "meta_test_%d.fa" % n))
    counts = orc.count_reads(reads, k)
    want = orc.component_cutter(counts, k, b1, b2)
    assert len(want) >= 2
    with m.KmerCounter(k) as ctx, m.KmerSet.load(ctx, [orc.kmers_bin(counts, 0, k)], 0) as ks:
        assert ks.components(b1, b2) == want
        assert ks.components(10 ** 9, 10 ** 9 + 1) == []            # everything is "small"
        assert ks.components(b1, b2) == want                        # repeatable on the same map
    with m.KmerCounter(k) as ctx, m.KmerSet(ctx) as empty:
        assert empty.components(1, 10) == []
