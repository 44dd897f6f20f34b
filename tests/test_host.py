"""CPU tests of the host side of libmfkc (parsers, naming, writers, generator) and of the C ABI
surface.  No GPU: no compute entry point is called."""
import ctypes as C
import gzip
import os
import re

import numpy as np
import pytest

import metafast_b200 as m
from metafast_b200 import _abi
from oracle import oracle as orc
from tests import _oracle_c
from tests.conftest import INPUTS, ROOT, has_gpu


def test_abi_exports_every_declared_symbol(built):
    """include/mfkc.h is the contract: every function it declares is exported and bound."""
    hdr = open(os.path.join(ROOT, "include", "mfkc.h")).read()
    declared = set(re.findall(r"\b(mfkc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mfkc_cfg", "mfkc_synth_cfg"}
    lib = m.load()
    assert declared == set(_abi.SIGNATURES), declared ^ set(_abi.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.mfkc_abi_version() == 1
    assert C.sizeof(_abi.MfkcCfg) == 88 and C.sizeof(_abi.SynthCfg) == 80


def test_no_cpu_fallback(built):
    if has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(m.MfkcError) as e:
        m.KmerCounter(31)
    assert e.value.code == _abi.E_CUDA and "no CPU fallback" in e.value.msg


def test_product_does_not_touch_the_oracle():
    """the product path must never import, link or execute anything under oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "metafast_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("oracle/ref_cpu.c", "").replace("the oracle", "") or f == "counter.py", f


def _write(p, text, gz=False):
    if gz:
        with gzip.open(p, "wb") as f:
            f.write(text.encode("latin-1"))
    else:
        with open(p, "wb") as f:
            f.write(text.encode("latin-1"))


FASTA_TRICKY = (">r1 multi-line\nACGTAC\nGTTTGA\n;comment line\nGGGCCC\r\n>r2 with N\nACGTNACGT\n>r3 lower\nacgtacgtaa\n"
                ">empty\n>r4\n\nAC\n\nGT\n>r5 n\nacgtn\n>last no newline\nTTTTGGGG")


@pytest.mark.parametrize("gz", [False, True])
def test_fasta_parser_rules(built, tmp_path, gz):
    p = str(tmp_path / ("x.fa.gz" if gz else "x.fasta"))
    _write(p, FASTA_TRICKY, gz)
    want = ["ACGTACGTTTGA", "GGGCCC", "acgtacgtaa", "ACGT", "TTTTGGGG"]
    assert orc.parse_reads(p) == want
    assert m.read_file_reads(p) == want
    b, o = _oracle_c.parse_file(p)
    assert [bytes(b[int(o[i]):int(o[i + 1])]).decode() for i in range(len(o) - 1)] == want
    assert m.reader_name(p) == "x" == orc.library_name(p)


def _fastq(records):
    return "".join("@%s\n%s\n+%s\n%s\n" % (n, s, n, q) for n, s, q in records)


def test_fastq_parser_rules(built, tmp_path):
    recs = [("a", "ACGTACGT", "IIIIIII5"),       # '5' < 64 -> the file is Sanger
            ("b", "ACGTNCGT", "IIII!III"),       # N -> dropped
            ("c", "ACGTACGT", "IIII!III"),       # phred 0 under a real base -> dropped
            ("d", "acgtacgt", "IIIIIIII"),
            ("e", "ACGT.CGT", "IIIIIIII"),       # '.' -> dropped
            ("f", "ACGTACGT", "IIIaIIII"),       # 'a' = phred 64 -> 6-bit alias of 0 -> dropped (DnaQ.java:140-150)
            ("g", "", ""),
            ("h", "TTTT", "@@@@")]               # '@' in Sanger = phred 31: kept
    text = _fastq(recs[:3]) + "\n\n" + _fastq(recs[3:])             # empty lines between records are skipped
    p = str(tmp_path / "s_R1.fastq")
    _write(p, text)
    want = ["ACGTACGT", "acgtacgt", "", "TTTT"]
    assert orc.parse_reads(p) == want
    assert m.read_file_reads(p) == want
    b, o = _oracle_c.parse_file(p)
    assert [bytes(b[int(o[i]):int(o[i + 1])]).decode() for i in range(len(o) - 1)] == want
    # Illumina (+64) file: every quality char >= 64 in the first 1000 reads; '@' is then phred 0
    ill = _fastq([("a", "ACGT", "hhhh"), ("b", "ACGT", "hh@h"), ("c", "GGGG", "efgh")])
    p2 = str(tmp_path / "ill.fq.gz")
    _write(p2, ill, gz=True)
    assert orc.parse_reads(p2) == ["ACGT", "GGGG"] == m.read_file_reads(p2)
    assert m.reader_name(p2) == "ill"
    # CRLF line ends
    p3 = str(tmp_path / "crlf.fastq")
    _write(p3, _fastq([("a", "ACGT", "II5I")]).replace("\n", "\r\n"))
    assert m.read_file_reads(p3) == ["ACGT"] == orc.parse_reads(p3)


def test_parser_errors(built, tmp_path):
    cases = {"bad_struct.fastq": "ACGT\nIIII\n", "len.fastq": "@a\nACGT\n+\nIII\n", "trunc.fastq": "@a\nACGT\n+\n",
             "badchar.fa": ">a\nACGTXACGT\n", "iupac.fa": ">a\nACGTRACGT\n", "qual.fastq": "@a\nACGT\n+\nII\x1fI\n"}
    for name, text in cases.items():
        p = str(tmp_path / name)
        _write(p, text)
        with pytest.raises(Exception):
            orc.parse_reads(p)
        with pytest.raises(m.MfkcError) as e:
            m.read_file_reads(p)
        assert e.value.code == _abi.E_FORMAT, name
    with pytest.raises(m.MfkcError):
        m.read_file_reads(str(tmp_path / "unknown.txt"))
    with pytest.raises(m.MfkcError):
        m.read_file_reads(str(tmp_path / "missing.fa"))


def test_reader_matches_oracle_on_fixtures(built):
    for f in ("meta_test_1.fa", "meta_test_2.fa", "meta_test_3.fa", "tinytest_A.fastq", "tinytest_B.fastq"):
        p = os.path.join(INPUTS, f)
        assert m.read_file_reads(p) == orc.parse_reads(p)
    # small batches exercise the carry-over of a read that does not fit
    got = []
    for b, o in m.read_file(os.path.join(INPUTS, "meta_test_2.fa"), batch_reads=7, batch_bases=1000):
        s = b.tobytes().decode()
        got += [s[int(o[i]):int(o[i + 1])] for i in range(len(o) - 1)]
    assert got == orc.parse_reads(os.path.join(INPUTS, "meta_test_2.fa"))


def test_sample_grouping():
    """src/tools/KmersCounterForManyFilesMain.java:73-108, KmersCounterMain.java:122-137"""
    files = ["/d/b_R2.fq.gz", "/d/a.fa", "/d/b_R1.fq.gz", "/d/c_r1.fastq", "/d/c_r2.fastq", "/d/z_R1.fq", "/d/x_R2.fq"]
    got = orc.group_samples(files)
    assert got == [("a", ["/d/a.fa"]), ("b", ["/d/b_R1.fq.gz", "/d/b_R2.fq.gz"]), ("c", ["/d/c_r1.fastq", "/d/c_r2.fastq"]),
                   ("x_R2", ["/d/x_R2.fq"]), ("z_R1", ["/d/z_R1.fq"])]


def test_stat_file_writer(built, tmp_path):
    counts = {1: 3, 2: 3, 3: 1, 9: 32767}
    hist = np.zeros(m.HIST_BINS, dtype=np.uint64)
    for c in counts.values():
        hist[c] += 1
    p = str(tmp_path / "x.stat.txt")
    m.write_stat_file(p, hist)
    assert open(p).read() == orc.stat_txt(counts) == "# k-mer frequency\tnumber of such k-mers\n1\t1\n3\t2\n32767\t1\n\n"


def test_synth_generator_is_deterministic(built):
    cfg = m.synth_cfg(total_genome_bp=500000, n_genomes=6)
    a = m.synth_reads_host(cfg, 0, 5000)
    b = m.synth_reads_host(cfg, 1000, 1000)
    assert (a[1000:2000] == b).all()                                  # counter-based: read i does not depend on the range
    assert set(np.unique(a)) <= set(b"ACGTN")
    other = m.synth_reads_host(m.synth_cfg(total_genome_bp=500000, n_genomes=6, sample=1), 0, 1000)
    assert not (other == a[:1000]).all()                              # per-sample abundances / reads
    # coverage makes k-mers repeat: far fewer distinct than instances
    keep = ~(a == ord("N")).any(axis=1)
    reads = [bytes(r).decode() for r in a[keep][:1500]]
    counts = orc.count_reads(reads, 31)
    assert len(counts) < 0.9 * sum(len(r) - 30 for r in reads)


def test_owner_shard_is_a_partition(built):
    lib = m.load()
    rng = np.random.default_rng(1)
    keys = [int(x) for x in rng.integers(0, 1 << 62, 2000)] + [0, 1, (1 << 62) - 1]
    for g in (1, 2, 3, 8):
        owners = [lib.mfkc_owner_shard(k, g) for k in keys]
        assert all(0 <= o < g for o in owners)
        if g > 1:
            assert len(set(owners)) == g and max(np.bincount(owners)) < 2.0 * len(keys) / g


def _read_all(path, threads, chunk=None, monkeypatch=None):
    monkeypatch.setenv("MFKC_READER_THREADS", str(threads))
    if chunk:
        monkeypatch.setenv("MFKC_READER_CHUNK", str(chunk))
    else:
        monkeypatch.delenv("MFKC_READER_CHUNK", raising=False)
    try:
        return ("ok", m.read_file_reads(path))
    except m.MfkcError as e:
        return ("error", e.code, e.msg)


def test_parallel_reader_equals_serial(built, tmp_path, monkeypatch):
    """the chunked multi-threaded ingest must hand out exactly what the serial parser does --
    kept reads in file order, and the same error for a malformed file -- whatever the chunk size."""
    import gzip
    rng = np.random.default_rng(3)
    seq = lambda n: "".join(rng.choice(list("ACGT"), n))
    fq, fa = [], []
    for i in range(400):
        s = seq(int(rng.integers(1, 120)))
        if i % 37 == 0:
            s = s[: len(s) // 2] + "N" + s[len(s) // 2 + 1:]
        q = "".join(rng.choice(list("#5?FI"), len(s)))
        fq.append("@r%d extra\n%s\n+\n%s\n" % (i, s, q) + ("\n" if i % 11 == 0 else ""))     # empty lines between records
        fa.append((">c%d\n" if i % 5 else ";c%d\n") % i + "\n".join(s[j:j + 30] for j in range(0, len(s), 30)) + "\n")
    cases = {
        "a.fastq": "".join(fq),
        "crlf.fastq": "".join(fq).replace("\n", "\r\n"),
        "cr.fastq": "".join(fq[:50]).replace("\n", "\r"),
        "notail.fastq": "".join(fq).rstrip("\n"),
        "trunc.fastq": "".join(fq)[: len("".join(fq)) // 2],
        "badchar.fastq": "".join(fq[:200]) + "@x\nACGTXACGT\n+\nIIIIIIIII\n" + "".join(fq[200:]),
        "badlen.fastq": "".join(fq[:100]) + "@x\nACGT\n+\nIII\n" + "".join(fq[100:]),
        "a.fa": "".join(fa),
        "lead.fa": "ACGTACGT\nACGT\n" + "".join(fa),                      # data before the first header is a record too
        "crlf.fa": "".join(fa).replace("\n", "\r\n"),
        "iupac.fa": "".join(fa[:100]) + ">bad\nACGTRYACGT\n" + "".join(fa[100:]),
        "empty.fa": "",
    }
    for name, text in cases.items():
        p = tmp_path / name
        p.write_bytes(text.encode())
        want = _read_all(str(p), 1, monkeypatch=monkeypatch)
        for threads, chunk in ((2, 64), (4, 1000), (3, None)):
            assert _read_all(str(p), threads, chunk, monkeypatch) == want, (name, threads, chunk)
        if name in ("a.fastq", "a.fa"):
            assert want[0] == "ok" and len(want[1]) > 300
            gz = tmp_path / (name + ".gz")
            with gzip.open(gz, "wb") as f:
                f.write(text.encode())
            assert _read_all(str(gz), 4, 777, monkeypatch) == want
        if name in ("trunc.fastq", "badchar.fastq", "badlen.fastq", "iupac.fa"):
            assert want[0] == "error"


def test_fastq_cuts_from_scouted_newline_counts(built, tmp_path, monkeypatch):
    """Plain FASTQ: scout threads count the newlines block by block, the producer derives the record cuts from the running
    count (mod 4) and hands over to the exact line machine at the first block that is not of the simple shape.  Same reads
    as the serial parser, whatever the number of scouts and the block size; quality lines may start with '@' or '+'."""
    rng = np.random.default_rng(21)
    seq = lambda n: "".join(rng.choice(list("ACGT"), n))
    recs = []
    for i in range(2500):
        s = seq(int(rng.integers(1, 140)))
        q = "".join(rng.choice(list("@+#5?FI"), len(s)))
        recs.append("@r%d\n%s\n+\n%s\n" % (i, s, q))
    text = "".join(recs)
    cases = {
        "clean.fastq": text,
        "notail.fastq": text.rstrip("\n"),
        "crlf_late.fastq": text + "".join(recs[:300]).replace("\n", "\r\n") + text,
        "blank_mid.fastq": "".join(recs[:1200]) + "\n\n" + "".join(recs[1200:]),
        "blank_first.fastq": "\n" + text,
        "trunc.fastq": text[: len(text) * 2 // 3],
    }
    for name, t in cases.items():
        p = tmp_path / name
        p.write_bytes(t.encode())
        monkeypatch.delenv("MFKC_READER_SCOUTS", raising=False)
        want = _read_all(str(p), 1, monkeypatch=monkeypatch)
        for scouts in (0, 1, 3):
            monkeypatch.setenv("MFKC_READER_SCOUTS", str(scouts))
            for threads, chunk in ((2, 256), (4, 5000), (3, 70000)):
                assert _read_all(str(p), threads, chunk, monkeypatch) == want, (name, scouts, threads, chunk)
    monkeypatch.delenv("MFKC_READER_SCOUTS", raising=False)


def test_record_longer_than_the_batch_buffer(built, tmp_path, monkeypatch):
    """The reference takes FASTA records of any length (FastaReader.java:54-108).  mfkc_reader_next keeps a read that
    does not fit the caller's buffer pending and tells its length; the same reads come out whatever the buffer, the
    chunk size and the number of threads, and a record that spans dozens of producer chunks is stepped over once."""
    import ctypes as C
    rng = np.random.default_rng(11)
    seq = lambda n: "".join(rng.choice(list("ACGT"), n))
    long1, long2 = seq(300_000), seq(90_001)
    recs = [seq(50), long1, seq(70), seq(1), long2]
    text = "".join(">r%d some; text > here\n" % i + "\n".join(r[j:j + 60] for j in range(0, len(r), 60)) + "\n" for i, r in enumerate(recs))
    fa = tmp_path / "long.fa"
    fa.write_bytes(text.encode())
    fq = tmp_path / "long.fastq"
    fq.write_bytes("".join("@r%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)) for i, r in enumerate(recs)).encode())
    lib = m.load()
    for path in (fa, fq):
        for threads, chunk in ((1, None), (3, 4096), (4, None)):
            monkeypatch.setenv("MFKC_READER_THREADS", str(threads))
            if chunk:
                monkeypatch.setenv("MFKC_READER_CHUNK", str(chunk))
            else:
                monkeypatch.delenv("MFKC_READER_CHUNK", raising=False)
            got = []
            for b, o in m.read_file(str(path), batch_reads=8, batch_bases=1000):
                s = b.tobytes().decode()
                got += [s[int(o[i]):int(o[i + 1])] for i in range(len(o) - 1)]
            assert got == recs, (path.name, threads, chunk)
            # the raw protocol: error, pending length, retry
            h, err, n = C.c_void_p(), C.create_string_buffer(512), C.c_uint32()
            assert lib.mfkc_reader_open(str(path).encode(), C.byref(h), err, 512) == 0
            small = np.empty(100, dtype=np.uint8); offs = np.empty(9, dtype=np.uint64); pend = C.c_uint64()
            vp = lambda a: a.ctypes.data_as(C.c_void_p)
            assert lib.mfkc_reader_next(h, vp(small), small.nbytes, vp(offs), 8, C.byref(n)) == 0 and n.value == 1     # r0 alone fits
            assert lib.mfkc_reader_next(h, vp(small), small.nbytes, vp(offs), 8, C.byref(n)) == _abi.E_BADARG
            assert b"longer than the batch buffer" in lib.mfkc_reader_error(h)
            assert lib.mfkc_reader_pending_bases(h, C.byref(pend)) == 0 and pend.value == len(long1)
            big = np.empty(pend.value, dtype=np.uint8)
            assert lib.mfkc_reader_next(h, vp(big), big.nbytes, vp(offs), 8, C.byref(n)) == 0 and n.value == 1
            assert big.tobytes().decode() == long1
            lib.mfkc_reader_close(h)


# ---------------------------------------------------------------- the main caller's remaining stages (matrix-builder)
def _emulated_components(hm, k, b1, b2, grid=1):
    """tests/emu/cc_emu.cpp: the kernels of metafast_b200/csrc/components.cuh compiled for the host"""
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libcc_emu.so"))
    keys = np.array(sorted(hm), dtype=np.uint64)
    vals = np.array([hm[int(x)] & 0xFFFF for x in keys], dtype=np.uint32)
    nc, nk, lv, res = C.c_uint64(), C.c_uint64(), C.c_int(), C.c_void_p()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.cc_emu_components(vp(keys), vp(vals), C.c_uint64(len(keys)), C.c_int(k), C.c_longlong(b1), C.c_longlong(b2),
                          C.byref(nc), C.byref(nk), C.byref(lv), C.byref(res), C.c_int(grid))
    off = np.zeros(nc.value + 1, dtype=np.uint64)
    ks = np.zeros(max(nk.value, 1), dtype=np.int64)
    w = np.zeros(max(nc.value, 1), dtype=np.int64)
    t = np.zeros(max(nc.value, 1), dtype=np.int32)
    lib.cc_emu_fetch(res, vp(off), vp(ks), vp(w), vp(t))
    return [(int(w[i]), [int(x) for x in ks[int(off[i]):int(off[i + 1])]], int(t[i])) for i in range(nc.value)], lv.value


def test_component_kernels_emulated_on_host(built):
    """component-cutter's graph half: the CUDA kernels (union-find levels, classification) and the host grouping, run as
    one emulated thread on the CPU, against the restatement of ComponentsBuilder -- size window, re-split levels, order"""
    reads = orc.parse_reads(os.path.join(INPUTS, "meta_test_2.fa"))
    for k, b1, b2, min_levels in ((31, 50, 800, 3), (21, 1, 100, 5), (11, 30, 2000, 5), (5, 1, 20, 100)):
        hm = orc.count_reads(reads, k)
        want = orc.component_cutter(hm, k, b1, b2)
        got, levels = _emulated_components(hm, k, b1, b2)
        assert got == want and len(want) > 5
        assert levels >= min_levels                                   # big components were re-split level after level
        for rep in range(3):                                          # 8 concurrent emulated blocks: the union-find under contention
            assert _emulated_components(hm, k, b1, b2, grid=8)[0] == want
    assert _emulated_components({}, 31, 1, 10) == ([], 0)
    # values <= 0 are no vertices (`getValue() > 0`); key 0 (poly-A) is its own neighbour
    assert _emulated_components({0: 3, 5: 0xFFFE}, 3, 1, 10)[0] == [(3, [0], 1)] == orc.component_cutter({0: 3}, 3, 1, 10)


CLI = os.path.join(ROOT, "metafast_b200", "bin", "mfkc_cli")
GOLDEN_VECS = {"meta_test_1": [41935, 38354, 20375, 14211], "meta_test_2": [20208, 0, 0, 11337],
               "meta_test_3": [6517, 34484, 20359, 749]}            # tests/test_oracle.py::test_reference_matrix_golden


def _cli(*args):
    import subprocess
    r = subprocess.run([CLI] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    return r


def test_dist_matrix_and_renumbering_reproduce_the_reference_file(built, tmp_path):
    """dist-matrix-calculator + heatmap-maker's renumbering (host tools of mfkc_cli, no GPU involved) on the feature vectors
    of the three test samples: test_data/meta_test_matrix.txt byte for byte with --output-format %s, README.md:96-99 with
    the default %.4f"""
    vfiles = []
    for name, vec in GOLDEN_VECS.items():
        (tmp_path / (name + ".vec")).write_text(orc.vec_text(vec))
        vfiles.append(tmp_path / (name + ".vec"))
    r = _cli("-t", "dist-matrix-calculator", "--features", *vfiles, "--matrix-file", tmp_path / "m.txt", "--output-format", "%s")
    assert r.stdout.split() == [str(tmp_path / "m.txt")]
    r = _cli("-t", "heatmap-maker", "-i", tmp_path / "m.txt", "--output-format", "%s")
    assert r.stdout.split() == [str(tmp_path / "m_renumbered.txt")]
    golden = open(os.path.join(ROOT, "tests", "golden", "meta_test_matrix.txt")).read()
    assert (tmp_path / "m_renumbered.txt").read_text() == golden
    r = _cli("-t", "dist-matrix-calculator", "--features", *vfiles, "-w", tmp_path / "w")
    (out,) = r.stdout.split()
    assert re.fullmatch(r".*/w/dist_matrix_\d{4}\.\d\d\.\d\d_\d\d\.\d\d\.\d\d_original_order\.txt", out)
    assert open(out).read() == ("#\tmeta_test_1\tmeta_test_2\tmeta_test_3\nmeta_test_1\t0.0000\t0.5691\t0.2981\n"
                                "meta_test_2\t0.5691\t0.0000\t0.8448\nmeta_test_3\t0.2981\t0.8448\t0.0000\n")
    _cli("-t", "dist-matrix-calculator", "--features", *vfiles, "--matrix-file", tmp_path / "nn.txt", "-wn")
    assert (tmp_path / "nn.txt").read_text().splitlines()[0] == "0.0000\t0.5691\t0.2981"


def test_renumbering_and_java_number_format(built, tmp_path):
    """average-linkage order and java.util.Formatter's HALF_UP rounding of the shortest digits, against the restatement"""
    rng = np.random.default_rng(8)
    for n in (1, 2, 5, 9):
        a = rng.random((n, n))
        mat = (a + a.T) / 2
        np.fill_diagonal(mat, 0.0)
        if n >= 5:
            mat[0, 1] = mat[1, 0] = 0.125            # "%.2f": Java 0.13, C 0.12
            mat[2, 3] = mat[3, 2] = 0.99995          # carries into the integer part at %.4f
            mat[0, 4] = mat[4, 0] = 1.5e-7
            mat[1, 4] = mat[4, 1] = 12345.678
        names = ["s%d" % i for i in range(n)]
        src = tmp_path / ("in%d.txt" % n)
        src.write_text(orc.matrix_txt(mat.tolist(), names, None, "%s"))
        perm = orc.heatmap_order(mat.tolist())
        for fmt in ("%s", "%.4f", "%.2f", "%.0f"):
            out = tmp_path / ("out%d.txt" % n)
            _cli("-t", "heatmap-maker", "-i", src, "--newMatrix-file", out, "--output-format", fmt)
            want = orc.matrix_txt(mat.tolist(), names, perm, fmt)
            assert out.read_text() == want, (n, fmt)
    assert orc.java_format_fixed(0.125, 2) == "0.13" and orc.java_format_fixed(0.99995, 4) == "1.0000"


# ---------------------------------------------------------------- gzip decoder of the host ingest path (fast_inflate.h)
def _fast_inflate(data: bytes, piece: int = 1 << 20, cap: int = 1 << 26):
    import ctypes as C
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libinflate_harness.so"))
    lib.fi_inflate.restype = C.c_long
    lib.fi_inflate.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64)]
    out = np.zeros(cap, dtype=np.uint8)
    err = C.create_string_buffer(256)
    members = C.c_uint64()
    n = lib.fi_inflate(data, len(data), out.ctypes.data_as(C.c_void_p), cap, piece, err, 256, C.byref(members))
    if n < 0:
        return None, err.value.decode(), members.value
    return out[:n].tobytes(), "", members.value


def _zlib_gunzip(data: bytes):
    """concatenated members like gzread; None when zlib rejects the stream"""
    import zlib
    out, pos = b"", 0
    while pos < len(data):
        if data[pos:pos + 2] != b"\x1f\x8b":
            return out if pos else None                   # trailing garbage is ignored
        d = zlib.decompressobj(31)
        try:
            out += d.decompress(data[pos:])
        except zlib.error:
            return None
        if not d.eof:
            return None
        pos = len(data) - len(d.unused_data)
    return out


@pytest.mark.timeout(600)
def test_fast_inflate_matches_zlib(built):
    """every block type, code shape and member layout zlib can produce, read in pieces of several sizes"""
    import zlib
    rng = np.random.default_rng(21)

    def gz(raw, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=31, memlevel=8):
        c = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
        return c.compress(raw) + c.flush()

    texts = {
        "empty": b"", "one": b"A", "dna": bytes(rng.choice(list(b"ACGT"), 300000).astype(np.uint8)),
        "random": bytes(rng.integers(0, 256, 200000, dtype=np.uint8)), "zeros": bytes(100000),
        "fastq": b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(list(b"ACGTN"), 100, p=[.249, .249, .249, .249, .004]).astype(np.uint8)),
                                                  bytes(rng.choice(list(b"#5?FI"), 100).astype(np.uint8))) for i in range(3000)),
        "period3": b"ACG" * 50000, "period5": b"ACGTT" * 30000,
    }
    n_cases = 0
    for name, raw in texts.items():
        for level in (0, 1, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                comp = gz(raw, level, strategy, memlevel=1 if level == 1 else 8)     # memLevel 1: many small blocks
                for piece in (1 << 20, 4099, 1):
                    if piece == 1 and len(raw) > 1000:
                        continue
                    got, err, members = _fast_inflate(comp, piece)
                    assert got == raw, (name, level, strategy, piece, err)
                    assert members == 1
                    n_cases += 1
    # concatenated members (bgzip / pigz -i / cat a.gz b.gz), an empty member in between, trailing garbage
    multi = gz(texts["dna"]) + gz(b"") + gz(texts["fastq"], 1) + gz(texts["zeros"], 9)
    got, err, members = _fast_inflate(multi, 65536)
    assert got == texts["dna"] + texts["fastq"] + texts["zeros"] and members == 4
    assert _fast_inflate(multi + b"\x00garbage", 65536)[0] == got
    # header fields: FEXTRA + FNAME + FCOMMENT + FHCRC
    body = gz(texts["fastq"])[10:]
    hdr = b"\x1f\x8b\x08\x1e\0\0\0\0\0\x03" + b"\x06\x00BC\x02\x00\x12\x34" + b"name.fq\0" + b"a comment\0"
    hcrc = (zlib.crc32(hdr) & 0xFFFF).to_bytes(2, "little")                 # FHCRC = low 16 bits of the header's CRC-32 (RFC 1952)
    assert _fast_inflate(hdr + hcrc + body)[0] == texts["fastq"]
    got_bad, err_bad, _ = _fast_inflate(hdr + bytes([hcrc[0] ^ 1, hcrc[1]]) + body)
    assert got_bad is None and "header CRC" in err_bad
    # a member must not reach into the previous member's text
    assert n_cases > 100


@pytest.mark.timeout(600)
def test_fast_inflate_rejects_what_zlib_rejects(built):
    """truncated and bit-flipped streams: an error (with zlib's wording where it has one), never silent garbage"""
    import zlib
    rng = np.random.default_rng(22)
    raw = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(list(b"ACGT"), 80).astype(np.uint8)), b"I" * 80) for i in range(800))
    comp = zlib.compressobj(6, zlib.DEFLATED, 31)
    comp = comp.compress(raw) + comp.flush()
    assert _fast_inflate(comp)[0] == raw
    assert _fast_inflate(b"not gzip at all")[1] == "not a gzip file"
    assert _fast_inflate(comp[:-4] + b"\0\0\0\0")[1] == "incorrect length check"
    assert _fast_inflate(comp[:-8] + b"\0\0\0\0" + comp[-4:])[1] == "incorrect data check"
    agree = 0
    for t in range(400):
        bad = bytearray(comp)
        if t % 4 == 0:
            bad = bad[:int(rng.integers(0, len(bad)))]
        else:
            bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        want = _zlib_gunzip(bytes(bad))
        got, err, _ = _fast_inflate(bytes(bad))
        if want is None:
            assert got is None and err, t
        else:
            assert got == want, t                        # a flip in MTIME / OS / a name: both accept
        agree += 1
    assert agree == 400


def test_crc32_clmul_matches_zlib(built):
    import ctypes as C
    import zlib
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libinflate_harness.so"))
    lib.fi_crc32.restype = C.c_uint32
    lib.fi_crc32.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
    rng = np.random.default_rng(23)
    for n in list(range(0, 200)) + [255, 256, 257, 1000, 4096, 65535, 1 << 20, (1 << 20) + 17]:
        data = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        start = int(rng.integers(0, 1 << 32)) if n % 2 else 0
        assert lib.fi_crc32(start, data, n) == zlib.crc32(data, start), n


@pytest.mark.timeout(600)
def test_gz_reader_fast_and_zlib_paths_agree(built, tmp_path, monkeypatch):
    """the reader over .gz files: FastInflate (default) and MFKC_INFLATE=zlib hand out the same reads; a corrupt file is an
    error in both; a '.gz' name on plain text is read transparently, like gzopen does"""
    rng = np.random.default_rng(24)
    recs = []
    for i in range(5000):
        n = int(rng.integers(30, 200))
        seq = bytes(rng.choice(list(b"ACGTN"), n, p=[.2495, .2495, .2495, .2495, .002]).astype(np.uint8))
        recs.append(b"@r%d\n%s\n+\n%s\n" % (i, seq, bytes(rng.choice(list(b"#5?FI"), n).astype(np.uint8))))
    text = b"".join(recs)
    one = tmp_path / "one.fastq.gz"
    with gzip.open(one, "wb") as f:
        f.write(text)
    multi = tmp_path / "multi.fastq.gz"
    with open(multi, "wb") as f:
        for part in (text[:len(text) // 3], b"", text[len(text) // 3:]):
            f.write(gzip.compress(part, 1))
    plain_named_gz = tmp_path / "plain.fastq.gz"
    plain_named_gz.write_bytes(text)
    want = orc.parse_reads(str(one))
    monkeypatch.setenv("MFKC_READER_CHUNK", "70000")
    for path in (one, multi, plain_named_gz):
        for mode in ("", "zlib"):
            if mode:
                monkeypatch.setenv("MFKC_INFLATE", mode)
            else:
                monkeypatch.delenv("MFKC_INFLATE", raising=False)
            assert m.read_file_reads(str(path)) == want, (path, mode)
    bad = bytearray(one.read_bytes())
    bad[len(bad) // 2] ^= 0x10
    corrupt = tmp_path / "corrupt.fastq.gz"
    corrupt.write_bytes(bytes(bad))
    for mode in ("", "zlib"):
        if mode:
            monkeypatch.setenv("MFKC_INFLATE", mode)
        else:
            monkeypatch.delenv("MFKC_INFLATE", raising=False)
        with pytest.raises(m.MfkcError):
            m.read_file_reads(str(corrupt))


def _parallel_inflate(data: bytes, threads: int, segment: int, piece: int = 1 << 20, cap: int = 1 << 27):
    import ctypes as C
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libinflate_harness.so"))
    lib.pi_inflate.restype = C.c_long
    lib.pi_inflate.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_char_p, C.c_size_t]
    out = np.zeros(cap, dtype=np.uint8)
    err = C.create_string_buffer(256)
    n = lib.pi_inflate(data, len(data), out.ctypes.data_as(C.c_void_p), cap, piece, threads, segment, err, 256)
    if n == -3:
        return "declined", "", 0
    if n < 0:
        return None, err.value.decode(), 0
    return out[:n].tobytes(), "", n


def _fastq_text(rng, n_reads, read_len=100):
    """FASTQ with random bases, varied qualities and repeated reads (long-range copies, so that markers of the unknown
    window travel far)"""
    bases = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), (n_reads, read_len))
    dup = rng.integers(0, n_reads, n_reads // 5)
    bases[dup] = bases[(dup * 7 + 1) % n_reads]
    quals = rng.choice(np.frombuffer(b"#5?FIII", dtype=np.uint8), (n_reads, read_len))
    out = bytearray()
    for i in range(n_reads):
        out += b"@read_%d\n" % i + bases[i].tobytes() + b"\n+\n" + quals[i].tobytes() + b"\n"
    return bytes(out)


@pytest.mark.timeout(600)
def test_parallel_inflate_is_exact(built):
    """one gzip stream decoded by several threads (csrc/parallel_inflate.h): identical to zlib for every thread count and
    segment size -- tiny segments put dozens of speculative starts, meetings and window hand-overs into a few megabytes"""
    import zlib
    rng = np.random.default_rng(31)
    text = _fastq_text(rng, 60000)                                            # ~13 MB

    def gz(raw, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, memlevel=8):
        c = zlib.compressobj(level, zlib.DEFLATED, 31, memlevel, strategy)
        return c.compress(raw) + c.flush()

    for level, memlevel in ((6, 8), (1, 8), (9, 9), (6, 1)):
        comp = gz(text, level, memlevel=memlevel)
        for threads, segment in ((2, 1 << 20), (3, 1 << 18), (8, 1 << 16), (5, 40000), (8, 1 << 14)):
            got, err, _ = _parallel_inflate(comp, threads, segment, piece=(1 << 20) + 13)
            assert got == text, (level, memlevel, threads, segment, err)
    assert _parallel_inflate(gz(text[:100000]), 4, 1 << 20)[0] == "declined"   # too small: the caller uses FastInflate
    # no dynamic blocks to start from (fixed Huffman codes; stored blocks of random bytes): the sample probes of open() find
    # nothing and the decoder declines -- one FastInflate is faster than one run through the 16-bit path
    assert _parallel_inflate(gz(text[:3000000], 6, zlib.Z_FIXED), 4, 1 << 16)[0] == "declined"
    noise = bytes(rng.integers(0, 256, 3000000, dtype=np.uint8))
    assert _parallel_inflate(gz(noise), 4, 1 << 16)[0] == "declined"
    # one very long run: 50 MB without a single dynamic block in the middle of a text stream.  Its pieces are resolved and
    # handed over while the run is still decoding, and the decoder waits for the consumer instead of piling them up
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = [c.compress(text) + c.flush(zlib.Z_FULL_FLUSH)]
    c2 = zlib.compressobj(1, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)          # raw deflate, fixed codes, spliced in
    fixed = c2.compress(text * 4) + c2.flush(zlib.Z_FULL_FLUSH)
    long_text = text + text * 4 + text
    tail = c.compress(text) + c.flush()
    import struct
    spliced = parts[0] + fixed + tail[:-8] + struct.pack("<II", zlib.crc32(long_text), len(long_text) & 0xFFFFFFFF)
    assert _zlib_gunzip(spliced) == long_text
    assert _parallel_inflate(spliced, 4, 1 << 20, piece=1 << 16)[0] == long_text
    # text, then binary, then text again; full flushes in between (empty stored blocks, byte-aligned block starts)
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    mixed = c.compress(text[:4000000]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(noise[:500000]) + c.flush(zlib.Z_SYNC_FLUSH) + \
        c.compress(text[4000000:9000000]) + c.flush()
    assert _parallel_inflate(mixed, 6, 1 << 15)[0] == text[:4000000] + noise[:500000] + text[4000000:9000000]
    # several members: parallel for the first, FastInflate for the rest; trailing garbage ignored
    multi = gz(text[:8000000]) + gz(b"") + gz(text[8000000:], 1) + gz(noise[:1000])
    assert _parallel_inflate(multi, 4, 1 << 17)[0] == text + noise[:1000]
    assert _parallel_inflate(multi + b"\0\0\0trailing", 4, 1 << 17)[0] == text + noise[:1000]


@pytest.mark.timeout(600)
def test_parallel_inflate_reports_corruption(built):
    """flipped bits and truncation in a big stream: an error (or, where zlib accepts the stream, zlib's bytes) -- never
    silently different text.  The CRC-32 of the member is combined from the pieces' CRCs and checked like zlib does."""
    import zlib
    rng = np.random.default_rng(32)
    text = _fastq_text(rng, 20000)
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    comp = c.compress(text) + c.flush()
    assert _parallel_inflate(comp, 4, 1 << 16)[0] == text
    assert _parallel_inflate(comp[:-8] + b"\1\2\3\4" + comp[-4:], 4, 1 << 16)[1] == "incorrect data check"
    assert _parallel_inflate(comp[:-4] + b"\1\2\3\4", 4, 1 << 16)[1] == "incorrect length check"
    for t in range(60):
        bad = bytearray(comp)
        if t % 5 == 0:
            bad = bad[:int(rng.integers(len(bad) // 2, len(bad)))]
        else:
            bad[int(rng.integers(20, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        want = _zlib_gunzip(bytes(bad))
        got, err, _ = _parallel_inflate(bytes(bad), 4, 1 << 16)
        if want is None:
            assert got is None and err, (t, err)
        else:
            assert got == want, t


@pytest.mark.timeout(600)
def test_gz_reader_parallel_inflate_path(built, tmp_path, monkeypatch):
    """a .gz big enough for the multi-threaded decoder (>= 4 MB compressed): the reader hands out exactly the reads of the
    zlib path and of the serial decoder, for several decoder thread counts"""
    rng = np.random.default_rng(33)
    text = _fastq_text(rng, 110000)
    path = tmp_path / "big.fastq.gz"
    with gzip.open(path, "wb", compresslevel=6) as f:
        f.write(text)
    assert os.path.getsize(path) > (4 << 20)

    def digest():
        import hashlib
        h = hashlib.sha256()
        n = 0
        for bases, offs in m.read_file(str(path), batch_reads=1 << 16, batch_bases=1 << 24):
            h.update(bases.tobytes()); h.update(np.diff(offs).tobytes()); n += len(offs) - 1
        return n, h.hexdigest()

    monkeypatch.setenv("MFKC_INFLATE", "zlib")
    want = digest()
    assert want[0] == 110000
    monkeypatch.setenv("MFKC_INFLATE", "serial")
    assert digest() == want
    monkeypatch.delenv("MFKC_INFLATE")
    for threads in ("2", "5", "8"):
        monkeypatch.setenv("MFKC_INFLATE_THREADS", threads)
        assert digest() == want
    # closing a reader in mid-stream (decoder threads, parse workers and the producer are all busy) must not hang
    lib = m.load()
    for batches in (0, 1, 3):
        h = C.c_void_p()
        err = C.create_string_buffer(256)
        assert lib.mfkc_reader_open(str(path).encode(), C.byref(h), err, 256) == 0
        bases = np.zeros(1 << 22, dtype=np.uint8)
        offs = np.zeros((1 << 14) + 1, dtype=np.uint64)
        n = C.c_uint32()
        for _ in range(batches):
            assert lib.mfkc_reader_next(h, bases.ctypes.data_as(C.c_void_p), bases.size, offs.ctypes.data_as(C.c_void_p), 1 << 14, C.byref(n)) == 0
            assert n.value == 1 << 14
        lib.mfkc_reader_close(h)
    # a flipped bit in the middle of the big file: an error, after the reads in front of it
    bad = bytearray(path.read_bytes())
    bad[len(bad) // 2] ^= 4
    path.write_bytes(bytes(bad))
    with pytest.raises(m.MfkcError):
        digest()


def _bgzf(raw: bytes, block: int = 65280, level: int = 6) -> bytes:
    """bgzip's container: gzip members with the 'BC' extra field = member size - 1, and the empty EOF member"""
    import struct
    import zlib
    out = bytearray()
    for pos in list(range(0, len(raw), block)) + [None]:
        part = b"" if pos is None else raw[pos:pos + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = c.compress(part) + c.flush()
        size = 12 + 6 + len(body) + 8
        out += b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, size - 1) + body + \
            struct.pack("<II", zlib.crc32(part), len(part))
    return bytes(out)


@pytest.mark.timeout(600)
def test_bgzf_inflate(built):
    """bgzip'ed input: groups of members decoded side by side, handed out in order; errors and non-BGZF tails"""
    import ctypes as C
    import zlib
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libinflate_harness.so"))
    lib.bz_inflate.restype = C.c_long
    lib.bz_inflate.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_char_p, C.c_size_t]

    def run(data, threads=4, group=1 << 16, piece=(1 << 18) + 7, cap=1 << 26):
        out = np.zeros(cap, dtype=np.uint8)
        err = C.create_string_buffer(256)
        n = lib.bz_inflate(data, len(data), out.ctypes.data_as(C.c_void_p), cap, piece, threads, group, err, 256)
        return ("declined" if n == -3 else None if n < 0 else out[:n].tobytes()), err.value.decode()

    rng = np.random.default_rng(41)
    text = _fastq_text(rng, 40000)
    comp = _bgzf(text)
    assert _zlib_gunzip(comp) == text                                   # the container is what zlib reads too
    for threads, group in ((2, 1 << 16), (4, 1 << 18), (8, 1 << 15), (3, 100)):
        assert run(comp, threads, group)[0] == text
    assert run(gzip.compress(text))[0] == "declined"                    # ordinary gzip: ParallelInflate's job
    # an ordinary member and garbage behind the BGZF part
    assert run(comp + gzip.compress(b"tail text\n"))[0] == text + b"tail text\n"
    assert run(comp + b"\0\1\2")[0] == text
    # a flipped bit inside a member body: that member's CRC (or its codes) fail
    bad = bytearray(comp)
    bad[len(bad) // 2] ^= 0x20
    got, err = run(bytes(bad))
    assert got is None and err
    # through the reader
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "reads.fastq.gz")
        big = _fastq_text(rng, 130000)
        open(path, "wb").write(_bgzf(big))
        assert os.path.getsize(path) > (8 << 20)
        got_reads = m.read_file_reads(path)
        assert len(got_reads) == 130000
        lines = big.split(b"\n")
        assert got_reads[0] == lines[1].decode() and got_reads[-1] == lines[-4].decode()


def test_host_tools_error_paths(built, tmp_path):
    """dist-matrix-calculator / heatmap-maker: the reference's messages and exit code 1 (Tool.java:450-462)"""
    import subprocess

    def fails(*args):
        r = subprocess.run([CLI] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 1
        return r.stderr

    assert "Missing mandatory parameter --features" in fails("-t", "dist-matrix-calculator")
    assert "Failed to read features from" in fails("-t", "dist-matrix-calculator", "--features", tmp_path / "nope.vec")
    assert "Can't read matrix file" in fails("-t", "heatmap-maker", "-i", tmp_path / "nope.txt")
    (tmp_path / "empty.txt").write_text("")
    assert "No data to read in matrix file" in fails("-t", "heatmap-maker", "-i", tmp_path / "empty.txt")
    (tmp_path / "wide.txt").write_text("#\ta\tb\tc\na\t0\t1\t2\n")
    assert "columns' number > rows' number" in fails("-t", "heatmap-maker", "-i", tmp_path / "wide.txt")
    (tmp_path / "ragged.txt").write_text("#\ta\tb\na\t0\t1\nb\t1\n")
    assert "columns' number is different for different rows" in fails("-t", "heatmap-maker", "-i", tmp_path / "ragged.txt")
    (tmp_path / "long.txt").write_text("#\ta\na\t0\nb\t1\n")
    assert "too much rows" in fails("-t", "heatmap-maker", "-i", tmp_path / "long.txt")
    (tmp_path / "ok.txt").write_text("0.0\t0.5\n0.5\t0.0\n")
    assert "only %.<N>f, %f and %s are supported" in fails("-t", "heatmap-maker", "-i", tmp_path / "ok.txt", "--output-format", "%e")
    assert "is outside the path this build replaces" in fails("-t", "view")
    assert "Unrecognized option" in fails("-t", "heatmap-maker", "--nope")
    # without names: "<i> library" labels are made up for the picture only, the renumbered matrix has names
    r = subprocess.run([CLI, "-t", "heatmap-maker", "-i", str(tmp_path / "ok.txt")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0
    assert (tmp_path / "ok_renumbered.txt").read_text() == "#\t1 library\t2 library\n1 library\t0.0000\t0.5000\n2 library\t0.5000\t0.0000\n"


def test_compare_with_reference_tool(tmp_path):
    """tools/compare_with_reference.py: a reference-style tree (records in hash-map order, components and sequences in thread
    order) against a tree in this repository's order compares equal; a changed count does not"""
    import subprocess
    import sys
    rng = np.random.default_rng(51)
    reads = orc.parse_reads(os.path.join(INPUTS, "meta_test_2.fa"))
    counts = orc.count_reads(reads, 31)
    rec = orc.kmers_bin(counts, 1, 31)
    ref, new = tmp_path / "ref", tmp_path / "new"
    for d in (ref, new):
        for sub in ("kmers", "stats", "matrices", "sequences"):
            os.makedirs(d / sub)
    recs = [rec[i:i + 10] for i in range(0, len(rec), 10)]
    perm = rng.permutation(len(recs))
    (ref / "kmers" / "s.kmers.bin").write_bytes(b"".join(recs[i] for i in perm))       # hash-map order
    (new / "kmers" / "s.kmers.bin").write_bytes(rec)
    for d in (ref, new):
        (d / "stats" / "s.stat.txt").write_text(orc.stat_txt(counts))
    comps = orc.component_cutter(orc.count_reads(reads, 21), 21, 50, 800)
    (new / "components.bin").write_bytes(orc.save_components([(w, k) for w, k, _ in comps], 31))
    (ref / "components.bin").write_bytes(orc.save_components([(w, list(rng.permutation(k))) for w, k, _ in reversed(comps)], 31))
    (ref / "sequences" / "s.seq.fasta").write_text(">1 x\nACGT\nAC\n>2 y\nTTTT\n")
    (new / "sequences" / "s.seq.fasta").write_text(">1 y\nTTTT\n>2 x\nACGTAC\n")
    (ref / "matrices" / "dist_matrix_a_original_order.txt").write_text("#\ta\tb\na\t0.0000\t0.5000\nb\t0.5000\t0.0000\n")
    (new / "matrices" / "dist_matrix_b_original_order.txt").write_text("#\ta\tb\na\t0.0000\t0.5000\nb\t0.5000\t0.0000\n")
    tool = os.path.join(ROOT, "tools", "compare_with_reference.py")
    r = subprocess.run([sys.executable, tool, str(ref), str(new)], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "5 compared, 0 differ" in r.stdout, r.stdout
    bad = bytearray(rec)
    bad[9] ^= 1
    (new / "kmers" / "s.kmers.bin").write_bytes(bytes(bad))
    r = subprocess.run([sys.executable, tool, str(ref), str(new)], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "DIFFERS  records (multiset)  kmers/s.kmers.bin" in r.stdout


def test_fastq_ingest_of_the_golden_samples(built, tmp_path):
    """the reference's test samples re-written as FASTQ (plain, .gz, bgzip): every reader path hands out exactly the reads of
    the FASTA files the golden matrix is pinned on -- so the FASTQ / gzip ingest feeds the pinned pipeline unchanged"""
    for n in (1, 2, 3):
        fa = os.path.join(INPUTS, "meta_test_%d.fa" % n)
        want = m.read_file_reads(fa)
        assert want == orc.parse_reads(fa) and len(want) in (1917, 540, 1080)
        text = "".join("@r%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)) for i, r in enumerate(want)).encode()
        fq = tmp_path / ("s%d.fastq" % n)
        fq.write_bytes(text)
        (tmp_path / ("g%d.fastq.gz" % n)).write_bytes(gzip.compress(text))
        (tmp_path / ("b%d.fastq.gz" % n)).write_bytes(_bgzf(text, block=5000))
        for name in ("s%d.fastq", "g%d.fastq.gz", "b%d.fastq.gz"):
            assert m.read_file_reads(str(tmp_path / (name % n))) == want, name


@pytest.mark.parametrize("record_size,n_parts,threads", [(10, 1, 1), (10, 2, 1), (10, 8, 0), (10, 11, 3), (18, 4, 0)])
def test_native_record_merge(built, record_size, n_parts, threads):
    """mfkc_merge_records (multi-GPU output merge): interleave of per-shard key-sorted streams, any thread count."""
    from metafast_b200.sharded import merge_sorted_records
    rng = np.random.default_rng(record_size * 100 + n_parts)
    kb = record_size - 2
    n = 300_000 if threads != 1 else 20_000
    keys = np.unique(rng.integers(0, 2 ** 62, n, dtype=np.uint64))
    be = np.zeros((len(keys), record_size), dtype=np.uint8)
    be[:, kb - 8:kb] = keys.astype(">u8").view(np.uint8).reshape(-1, 8)
    if kb == 16:
        be[:, 0:8] = (keys % np.uint64(5)).astype(">u8").view(np.uint8).reshape(-1, 8)     # high word: few values, order decided by both
    be[:, kb:] = rng.integers(0, 256, (len(keys), 2), dtype=np.uint8)
    order = np.lexsort(tuple(be[:, i] for i in range(kb - 1, -1, -1)))
    be = be[order]
    owner = rng.integers(0, n_parts, len(keys))
    owner[: len(keys) // 3] = 0                                  # uneven parts, and (n_parts > 1) some part may be empty
    if n_parts > 2:
        owner[owner == n_parts - 1] = 1
    parts = [be[owner == p].tobytes() for p in range(n_parts)]
    got = merge_sorted_records(parts, record_size, threads)
    assert got == be.tobytes()
    assert merge_sorted_records([b""] * n_parts, record_size) == b""


def test_native_record_merge_ties_and_extreme_keys(built):
    """equal keys (the shards never produce them) keep part order; the all-ones and the zero key are ordinary keys"""
    from metafast_b200.sharded import merge_sorted_records
    rec = lambda key, tag: int(key).to_bytes(8, "big") + bytes([0, tag])
    top = 2 ** 64 - 1
    a = [rec(0, 1), rec(5, 1), rec(5, 2), rec(top, 1)]
    b = [rec(0, 3), rec(5, 3), rec(7, 3), rec(top, 3), rec(top, 4)]
    c = [rec(6, 5)]
    want = [rec(0, 1), rec(0, 3), rec(5, 1), rec(5, 2), rec(5, 3), rec(6, 5), rec(7, 3), rec(top, 1), rec(top, 3), rec(top, 4)]
    for threads in (1, 0):
        assert merge_sorted_records([b"".join(a), b"".join(b), b"".join(c)], 10, threads) == b"".join(want)
        assert merge_sorted_records([b"".join(a), b"", b"".join(c)], 10, threads) == b"".join(sorted(a + c, key=lambda r: r[:8]))
