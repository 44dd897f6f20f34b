"""CPU tests: the oracle against the golden vectors / known answers, and the two independent
restatements (numpy + C) against each other.  No GPU."""
import gzip
import json
import os
import struct

import numpy as np
import pytest

from oracle import oracle as orc
from tests import _oracle_c
from tests.conftest import GOLDEN, INPUTS


@pytest.mark.parametrize("n", [1, 2, 3])
def test_config1_golden(built, n):
    """BASELINE.md section 3 known answers for test_data/meta_test_{1,2,3}.fa, k=31, -b 1."""
    gold = json.load(open(os.path.join(GOLDEN, "config1.json")))["meta_test_%d" % n]
    survey = {1: (1917, 115020, 17063, 16918, 169180, "e660671016d5d1db", 15),
              2: (540, 32400, 8176, 7321, 73210, "c90d8b0b553e2701", 13),
              3: (1080, 64800, 14042, 11351, 113510, "591d0e32821afd8d", 15)}[n]
    path = os.path.join(INPUTS, "meta_test_%d.fa" % n)
    res = orc.kmer_counter_many([path], 31, 1)
    (name, (rec, stat, counts)), = res.items()
    assert name == "meta_test_%d" % n
    reads = orc.parse_reads(path)
    got = (len(reads), sum(len(r) - 30 for r in reads), len(counts), len(rec) // 10, len(rec),
           orc.sha256_hex(rec)[:16], max(counts.values()))
    assert got == survey                                             # SURVEY.md 8c / BASELINE.md 3
    assert orc.sha256_hex(rec) == gold["sha256_sorted_records"]
    assert {str(c): v for c, v in orc.histogram(counts).items()} == gold["hist"]
    assert orc.sha256_hex(stat.encode()) == gold["stat_txt_sha256"]
    assert stat.startswith("# k-mer frequency\tnumber of such k-mers\n1\t") and stat.endswith("\n\n")
    # the C restatement (threads + striped maps) gives the same file, whatever -p is
    bases, offsets = _oracle_c.parse_file(path)
    for P in (1, 3, 8):
        c_rec, c_hist, c_distinct, st = _oracle_c.count(bases, offsets, 31, 1, P=P)
        assert c_rec == rec and c_distinct == len(counts)
        assert st == [len(reads), len(reads), 90 * len(reads), 90 * len(reads)]
    # reference iteration order differs from key order but holds the same multiset
    u_rec, _, _, _ = _oracle_c.count(bases, offsets, 31, 1, P=4, sort=False)
    assert u_rec != rec and orc.sorted_records(u_rec) == rec


def test_micro_known_answers():
    """A0 G1 C2 T3, first base most significant, key = min(fw, rc)."""
    k5 = lambda s: orc.canonical_kmers(s, 5)[0]
    assert k5("AAAAA") == 0 and k5("TTTTT") == 0
    assert k5("GGGGG") == 341 and k5("CCCCC") == 341
    assert k5("ACGTA") == 156 and k5("AACAT") == 35
    assert orc.kmer_to_string(156, 5) in ("ACGTA", "TACGT")
    for k in (1, 5, 21, 31, 32, 55, 63):                             # rolling rc == KmerUtils.reverseComplement
        rng = np.random.default_rng(k)
        s = "".join(rng.choice(list("ACGT"), 200))
        mask = (1 << (2 * k)) - 1
        for i, key in enumerate(orc.canonical_kmers(s, k)):
            fw = 0
            for ch in s[i:i + k]:
                fw = (fw << 2) | orc.code(ch)
            assert key == min(fw & mask, orc.reverse_complement(fw, k))


def test_tinytest_fastq():
    reads = orc.parse_reads(os.path.join(INPUTS, "tinytest_A.fastq"))
    assert reads == ["AACATAAGC", "GAAGCCAAC"]                       # '#' < 64 -> Sanger, phred 2: kept
    assert sorted(orc.count_reads(reads, 5)) == [26, 35, 104, 140, 193, 262, 416, 444, 501, 560]


@pytest.mark.parametrize("k", [1, 4, 13, 16, 17, 31])
def test_numpy_vs_scalar_vs_c(built, k):
    rng = np.random.default_rng(100 + k)
    reads = ["".join(rng.choice(list("ACGTacgt"), int(L))) for L in rng.integers(0, 120, 400)]
    reads += ["A" * 50] * 300 + ["ACGT" * 10]
    want = {}
    for r in reads:
        for key in orc.canonical_kmers(r, k):
            want[key] = min(want.get(key, 0) + 1, 32767)
    assert orc.count_reads(reads, k) == want
    bases = np.frombuffer("".join(reads).encode(), dtype=np.uint8).copy()
    offsets = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    for b in (0, 1, 3):
        rec, hist, distinct, st = _oracle_c.count(bases, offsets, k, b, P=3)
        assert rec == orc.kmers_bin(want, b, k) and distinct == len(want)
        assert {int(c): int(hist[c]) for c in np.nonzero(hist)[0]} == orc.histogram(want)
    rec, _, _, st = _oracle_c.count(bases, offsets, k, 0, min_len=60, P=2)
    assert rec == orc.kmers_bin(orc.count_reads(reads, k, 60), 0, k)
    assert tuple(st) == orc.read_stats(reads, 60)


def test_saturation_cpu(built):
    reads = ["ACGTTGCA"] * 40000
    counts = orc.count_reads(reads, 5)
    assert set(counts.values()) == {32767}
    bases = np.frombuffer("".join(reads).encode(), dtype=np.uint8).copy()
    offsets = (np.arange(len(reads) + 1) * 8).astype(np.uint64)
    rec, hist, _, _ = _oracle_c.count(bases, offsets, 5, 1, P=4)
    assert rec == orc.kmers_bin(counts, 1, 5) and hist[32767] == len(counts)


def test_records_roundtrip():
    counts = {0: 5, 7: 1, (1 << 62) - 1: 32767, 12345678901234: 2}
    rec = orc.kmers_bin(counts, 1)
    assert len(rec) == 30 and rec[:10] == struct.pack(">qh", 0, 5)
    assert orc.load_kmers_bin(rec) == [(0, 5), (12345678901234, 2), ((1 << 62) - 1, 32767)]
    assert orc.load_kmers([rec, rec], 0) == {0: 10, 12345678901234: 4, (1 << 62) - 1: 32767}
    comps = [(7, [1, 2, 3]), (0, []), (-1, [5])]
    assert orc.load_components(orc.save_components(comps)) == comps


def test_java_double_to_string():
    f = orc.java_double_to_string
    assert [f(x) for x in (0.0, 1.0, 0.5, 1 / 3, 2 / 3, 0.001, 5e-4, 1e7, 1234567.0, 0.1, float("nan"))] == \
        ["0.0", "1.0", "0.5", "0.3333333333333333", "0.6666666666666666", "0.001", "5.0E-4", "1.0E7", "1234567.0", "0.1", "NaN"]


def test_features_oracle_agreement(built):
    rng = np.random.default_rng(3)
    counts = orc.count_reads(orc.parse_reads(os.path.join(INPUTS, "meta_test_2.fa")), 31)
    keys = sorted(counts)
    comps = [[keys[int(i)] for i in rng.integers(0, len(keys), int(rng.integers(0, 40)))] + [int(rng.integers(0, 1 << 62))]
             for _ in range(30)] + [[]]
    rec = orc.kmers_bin(counts, 1)
    sel = orc.kmers_bin({k: c for k, c in counts.items() if k % 2}, 0)
    for thr, s in ((0, None), (4, None), (0, sel), (1, b"")):
        acc = orc.presence_for_kmers([k for c in comps for k in c], orc.load_kmers_bin(rec))
        wv, wb, wf, wc = orc.features([(0, c) for c in comps], acc, thr, None if s is None else orc.load_kmers([s], 0))
        cv, cf, cc = _oracle_c.features_kmers(comps, rec, s, thr)
        assert list(cv) == wv and list(cf) == wf and list(cc) == wc
    assert orc.breadth_text([0.5, float("nan")]) == "0.5\nNaN\n" and orc.vec_text([3, -1]) == "3\n-1\n"
    # Java's addAndBound(long,long) with a negative increment saturates (NumUtils.java:27-32 quirk)
    assert orc._java_add_and_bound64(5, -1) == (1 << 63) - 1 and orc._java_add_and_bound64(5, 7) == 12


def test_set_algebra_known_answers():
    """hand-derived from the cited lines: addAndBound saturates, put(get + x) wraps like a Java short, getWithZero reads
    a stored -1 as 0 (src/io/IOUtils.java:249-257, src/tools/UniqueKmersMultipleSamplesFinder.java:106-129,
    [itmo]/structures/map/Long2ShortHashMap.java:160-183)"""
    import struct
    rec = lambda k, v: struct.pack(">Qh", k, v)
    # loadKmers: 3 + 32767 saturates, freq 1 is not > threshold 1
    assert orc.load_kmers([rec(5, 3) + rec(5, 32767) + rec(9, 1)], 1) == {5: 32767}
    # unique-kmers-multi, b = 0: 32767 + 32767 = 65534 -> (short) -2; -2 + 1 = -1; getWithZero(-1) = 0, so + 5 gives 5 again
    files = [rec(5, 32767), rec(5, 32767), rec(5, 1), rec(5, 5)]
    size, out = orc.unique_kmers_multi(files, [], 0, 1, 4)
    assert size == 1 and out[4] == rec(5, 5) and out[1] == rec(5, 5)
    # three files only: the sum is -1, which is not > b -> nothing printed although the k-mer is in 3 samples
    assert orc.unique_kmers_multi(files[:3], [], 0, 1, 3)[1] == {1: b"", 2: b"", 3: b""}
    # a filter sample that holds the k-mer zeroes it
    assert orc.unique_kmers_multi([rec(7, 4), rec(7, 4) + rec(8, 9)], [rec(7, 2)], 1, 1, 2)[1] == {1: rec(8, 9), 2: b""}
    # kmers-filter keeps what the known samples contain more than max-thresh * files times
    assert orc.kmers_filter([rec(1, 5) + rec(2, 5) + rec(3, 1)], [rec(1, 2), rec(1, 2) + rec(2, 1)], 0, 1) == [(3, rec(1, 5))]
    # kmers-samples-counter: number of samples a k-mer is good in
    n, recs, stat = orc.kmers_samples_counter([rec(1, 5) + rec(2, 5), rec(1, 2), rec(3, 9) + rec(1, 1)], 1)
    assert n == 3 and recs == rec(1, 2) + rec(2, 1) + rec(3, 1)
    assert stat == "# k-mer frequency\tnumber of such k-mers\n1\t2\n2\t1\n\n"


def test_reference_matrix_golden(built):
    """PINS THE ORACLE AGAINST THE REFERENCE'S OWN FIXTURE.  test_data/meta_test_matrix.txt (copied verbatim to
    tests/golden/meta_test_matrix.txt) is the output the reference's authors checked in for
    `matrix-builder -k 31 -i test_data/meta_test_{1,2,3}.fa` (README.md:90-99; defaults -b 1, -l 100): the Bray-Curtis
    distances between the three samples' feature vectors, printed with all 16-17 digits.  Every stage of the path
    feeds it: ~37 000 k-mer counts (parser, 2-bit code, canonical form, saturating count), the `count > b` filter of
    .kmers.bin, seq-builder, the minSeqLen counting call of component-cutter, the component split, and the
    features-calculator sums.  The oracle's restatement of that pipeline must reproduce the three distances to the
    last bit."""
    paths = [os.path.join(INPUTS, "meta_test_%d.fa" % n) for n in (1, 2, 3)]
    gold = orc.load_matrix_txt(open(os.path.join(GOLDEN, "meta_test_matrix.txt")).read())
    assert len(gold) == 9
    names, matrix, mid = orc.matrix_builder(paths, k=31, b=1, min_seq_len=100)
    assert names == ["meta_test_1", "meta_test_2", "meta_test_3"]
    for i, a in enumerate(names):
        for j, b in enumerate(names):
            assert matrix[i][j] == gold[(a, b)], (a, b, matrix[i][j], gold[(a, b)])      # exact doubles
    assert orc.java_double_to_string(matrix[0][1]) == "0.5691162409506898"
    assert orc.java_double_to_string(matrix[0][2]) == "0.2981399448537721"
    assert orc.java_double_to_string(matrix[1][2]) == "0.8448331091037222"
    # and the file itself, byte for byte: heatmap-maker's average-linkage renumbering puts meta_test_3 next to meta_test_1
    perm = orc.heatmap_order(matrix)
    assert perm == [0, 2, 1]
    assert orc.matrix_txt(matrix, names, perm, "%s") == open(os.path.join(GOLDEN, "meta_test_matrix.txt")).read()
    assert orc.matrix_txt(matrix, names, None, "%.4f").splitlines()[1] == "meta_test_1\t0.0000\t0.5691\t0.2981"   # README.md:97
    # intermediates, for the record (they are what the GPU pipeline test compares stage by stage)
    assert [len(mid["sequences"][n]) for n in names] == [15, 29, 25]
    assert len(mid["sequence_kmers"]) == 17061
    assert [(len(keys), w, thr) for w, keys, thr in mid["components"]] == \
        [(6240, 12783, 1), (5713, 11265, 1), (3020, 5977, 1), (2088, 4260, 1)]
    assert mid["vectors"]["meta_test_2"] == [20208, 0, 0, 11337]
    # the C restatement (oracle/ref_cpu.c: threads, striped maps) in place of the Python stages it covers -- parser + count
    # + filtered emit per sample, the minSeqLen count over the sequences, the feature sums -- gives the same matrix
    c_recs = []
    for p in paths:
        bases, offsets = _oracle_c.parse_file(p)
        c_recs.append(_oracle_c.count(bases, offsets, 31, 1, P=4)[0])
    seq_reads = [s[0] for n in names for s in mid["sequences"][n]]
    sb = np.frombuffer("".join(seq_reads).encode(), dtype=np.uint8).copy()
    so = np.zeros(len(seq_reads) + 1, dtype=np.uint64)
    so[1:] = np.cumsum([len(r) for r in seq_reads])
    c_seq = dict(orc.load_kmers_bin(_oracle_c.count(sb, so, 31, 0, min_len=100, P=3)[0]))
    assert c_seq == mid["sequence_kmers"]
    comps = [keys for _w, keys, _thr in orc.component_cutter(c_seq, 31)]
    c_vecs = [[int(x) for x in _oracle_c.features_kmers(comps, rec, None, 0)[0]] for rec in c_recs]
    for i in range(3):
        for j in range(3):
            if i != j:
                assert orc.bray_curtis(c_vecs[i], c_vecs[j]) == gold[(names[i], names[j])]
    # the fixture discriminates: the off-by-one reading of the -b filter (count >= b) moves every distance
    keep = orc.kmers_bin
    try:
        orc.kmers_bin = lambda counts, threshold, k=31: keep(counts, threshold - 1, k)
        _, wrong, _ = orc.matrix_builder(paths, k=31, b=1, min_seq_len=100)
    finally:
        orc.kmers_bin = keep
    assert all(wrong[i][j] != matrix[i][j] for i in range(3) for j in range(3) if i != j)


def test_component_cutter_known_answers():
    """ComponentsBuilder on a hand-made graph: size limits b1/b2, and the re-split of a big component at threshold + 1"""
    k = 11
    a = "TGGCCAAAATGTGGTGGGGTCTGACTGATGTAATAGACCCCAAAAGGGCGTCCTTTCGTG"
    b = "TGGCTAGGTGCCCCGTATGCGGCCGGGCTCCTCAG"
    hm = orc.count_reads([a, a, b], k)                     # two separate paths; the k-mers of `a` have count 2
    keys_a, keys_b = set(orc.canonical_kmers(a, k)), set(orc.canonical_kmers(b, k))
    assert len(keys_a) == 50 and len(keys_b) == 25
    assert not any(n in keys_b for x in keys_a for n in orc.possible_neighbours(x, k))
    shape = lambda comps: [(set(c[1]), c[0], c[2]) for c in comps]
    assert shape(orc.component_cutter(hm, k, b1=1, b2=1000)) == [(keys_a, 100, 1), (keys_b, 25, 1)]
    assert shape(orc.component_cutter(hm, k, b1=26, b2=1000)) == [(keys_a, 100, 1)]           # size < b1 is dropped
    assert shape(orc.component_cutter(hm, k, b1=25, b2=50)) == [(keys_a, 100, 1), (keys_b, 25, 1)]   # size == b2 is kept
    # b2 = 49: `a` (50 k-mers, all of count 2) is big at threshold 1, still big at threshold 2, and has no k-mer of count
    # >= 3: it vanishes.  b2 = 24: `b` is big too and has no k-mer of count >= 2
    assert shape(orc.component_cutter(hm, k, b1=1, b2=49)) == [(keys_b, 25, 1)]
    assert orc.component_cutter(hm, k, b1=1, b2=24) == []
    # a path whose two ends were seen twice: too big as a whole, its ends survive the re-split at threshold 2
    hm2 = orc.count_reads([a, a[:30], a[40:], b], k)
    head, tail = set(orc.canonical_kmers(a[:30], k)), set(orc.canonical_kmers(a[40:], k))
    assert shape(orc.component_cutter(hm2, k, b1=1, b2=49)) == [(keys_b, 25, 1), (head, 40, 2), (tail, 20, 2)]
    assert shape(orc.component_cutter(hm2, k, b1=11, b2=49)) == [(keys_b, 25, 1), (head, 40, 2)]
    comps = orc.component_cutter(hm, k, b1=1, b2=1000)
    assert orc.components_stat_txt(comps).splitlines()[1:] == ["1\t50\t100\t1", "2\t25\t25\t1"]
    assert orc.load_components(orc.save_components([(w, keys) for w, keys, _ in comps], k), k) == \
        [(w, keys) for w, keys, _ in comps]
    assert orc.bray_curtis([1, 2, 3], [1, 2, 3]) == 0.0 and orc.bray_curtis([1, 0], [0, 1]) == 1.0
    assert orc.bray_curtis([3, 1], [1, 1]) == 2.0 / 6.0
