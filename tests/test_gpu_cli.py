"""GPU tests of the native host tool (mfkc_cli): MetaFast's own command lines end to end --
files in, .kmers.bin / .stat.txt / .vec / .breadth out -- against the oracle, byte for byte."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

import metafast_b200 as m
from oracle import oracle as orc
from tests.conftest import INPUTS, ROOT

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "metafast_b200", "bin", "mfkc_cli")


def run_cli(*args, ok=True):
    r = subprocess.run([CLI] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if ok:
        assert r.returncode == 0, r.stderr
    return r


@pytest.mark.parametrize("variant", ["hash", "sort", "direct", "table"])
def test_kmer_counter_many_config1(built, tmp_path, variant):
    """BASELINE config 1 through the CLI: default -b (1), default output locations."""
    files = [os.path.join(INPUTS, "meta_test_%d.fa" % n) for n in (3, 1, 2)]
    wd = tmp_path / "wd"
    r = run_cli("-t", "kmer-counter-many", "-k", 31, "-i", *files, "-w", wd, "--gpu-variant", variant)
    want = orc.kmer_counter_many(files, 31, 1)
    assert r.stdout.split() == [str(wd / "kmers" / ("meta_test_%d.kmers.bin" % n)) for n in (1, 2, 3)]
    for name, (rec, stat, counts) in want.items():
        assert open(wd / "kmers" / (name + ".kmers.bin"), "rb").read() == rec
        assert open(wd / "stats" / (name + ".stat.txt")).read() == stat
    assert "17'063 k-mers found, 16'918 (99.2%) of them is good (not erroneous)" in r.stderr


@pytest.mark.parametrize("gpus,mode", [(2, "shard"), (3, "shard"), (2, "samples"), (8, "auto")])
def test_kmer_counter_many_multi_gpu(built, tmp_path, gpus, mode):
    """--gpus G: hash-range sharded over G contexts with the native merge (shard), or one sample per context at a time
    (samples); with fewer samples than GPUs `auto` shards.  Logical GPUs (G contexts on the one device of the test box,
    same kernels and peer pointers as across devices): every file byte-equal to the one-GPU / oracle result."""
    files = [os.path.join(INPUTS, "meta_test_%d.fa" % n) for n in (3, 1, 2)]
    wd = tmp_path / "wd"
    env = dict(os.environ, MFKC_LOGICAL_GPUS="1")
    r = subprocess.run([CLI, "-t", "kmer-counter-many", "-k", "31", "-i", *files, "-w", str(wd), "--gpus", str(gpus), "--gpu-mode", mode],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    assert r.returncode == 0, r.stderr
    want = orc.kmer_counter_many(files, 31, 1)
    assert r.stdout.split() == [str(wd / "kmers" / ("meta_test_%d.kmers.bin" % n)) for n in (1, 2, 3)]
    for name, (rec, stat, counts) in want.items():
        assert open(wd / "kmers" / (name + ".kmers.bin"), "rb").read() == rec
        assert open(wd / "stats" / (name + ".stat.txt")).read() == stat
    assert "17'063 k-mers found, 16'918 (99.2%) of them is good (not erroneous)" in r.stderr


def test_paired_gz_fastq_and_naming(built, tmp_path):
    """_R1/_R2 pairing into one sample, gzip input, Sanger sniffing, N-read dropping."""
    for tag, sample in (("R1", 0), ("R2", 1)):
        fq = tmp_path / ("lib_%s.fastq" % tag)
        run_cli("gen-reads", fq, 3000, sample, 200000, 5)
        with open(fq, "rb") as f, gzip.open(str(fq) + ".gz", "wb") as g:
            shutil.copyfileobj(f, g)
        os.remove(fq)
    single = tmp_path / "other.fa"
    run_cli("gen-reads", single, 1000, 2, 200000, 5)
    files = [str(tmp_path / "lib_R2.fastq.gz"), str(single), str(tmp_path / "lib_R1.fastq.gz")]
    out, st = tmp_path / "o", tmp_path / "s"
    run_cli("-t", "kmer-counter-many", "-k", 23, "-b", 2, "-i", *files, "--output-dir", out, "--stats-dir", st)
    want = orc.kmer_counter_many(files, 23, 2)
    assert sorted(want) == ["lib", "other"]
    for name, (rec, stat, counts) in want.items():
        assert open(out / (name + ".kmers.bin"), "rb").read() == rec
        assert open(st / (name + ".stat.txt")).read() == stat
    # kmer-counter (single sample) naming: two unpaired files -> "<first>+"
    run_cli("-t", "kmer-counter", "-k", 23, "-i", str(single), files[0], "-w", tmp_path / "w2")
    reads = orc.parse_reads(str(single)) + orc.parse_reads(files[0])
    assert open(tmp_path / "w2" / "kmers" / "other+.kmers.bin", "rb").read() == orc.kmers_bin(orc.count_reads(reads, 23), 1, 23)


def test_cli_errors(built, tmp_path):
    f = os.path.join(INPUTS, "meta_test_2.fa")
    assert run_cli("-t", "kmer-counter-many", "-k", 32, "-i", f, ok=False).returncode == 1     # KmersCounterMain.java:70-73
    assert run_cli("-t", "kmer-counter-many", "-k", 0, "-i", f, ok=False).returncode == 1
    assert run_cli("-t", "kmer-counter-many", "-i", f, ok=False).returncode == 1               # -k is mandatory
    bad = tmp_path / "bad.fa"
    bad.write_text(">a\nACGTXXACGT\n")
    assert run_cli("-t", "kmer-counter-many", "-k", 5, "-i", bad, ok=False).returncode == 1
    assert run_cli("-t", "component-cutter", "-k", 5, ok=False).returncode == 1


def test_features_calculator_cli(built, tmp_path):
    rng = np.random.default_rng(17)
    files = [os.path.join(INPUTS, "meta_test_%d.fa" % n) for n in (1, 2, 3)]
    wd = tmp_path / "wd"
    run_cli("-t", "kmer-counter-many", "-k", 31, "-i", *files, "-w", wd)
    per_sample = orc.kmer_counter_many(files, 31, 1)
    all_keys = sorted(set().union(*[set(c) for _, _, c in per_sample.values()]))
    comps = []
    for _ in range(60):
        size = int(rng.integers(1, 400))
        comps.append((int(rng.integers(0, 1000)), [all_keys[int(i)] for i in rng.integers(0, len(all_keys), size)]))
    comps.append((0, []))
    comps.append((5, [int(x) for x in rng.integers(0, 1 << 62, 7)]))          # k-mers no sample contains
    cm = tmp_path / "components.bin"
    cm.write_bytes(orc.save_components(comps))
    kfiles = [str(wd / "kmers" / ("meta_test_%d.kmers.bin" % n)) for n in (1, 2, 3)]
    fw = tmp_path / "fw"
    r = run_cli("-t", "features-calculator", "-k", 31, "-cm", cm, "-ka", *kfiles, "-w", fw)
    for n in (1, 2, 3):
        rec = per_sample["meta_test_%d" % n][0]
        acc = orc.presence_for_kmers([k for _, c in comps for k in c], orc.load_kmers_bin(rec))
        vec, breadth, _, _ = orc.features(comps, acc, 0)
        assert open(fw / "vectors" / ("meta_test_%d.vec" % n)).read() == orc.vec_text(vec)
        assert open(fw / "vectors" / ("meta_test_%d.breadth" % n)).read() == orc.breadth_text(breadth)
    # --threshold, --selected and reads mode (-i)
    fw2 = tmp_path / "fw2"
    run_cli("-t", "features-calculator", "-k", 31, "-cm", cm, "-ka", kfiles[0], "-i", files[1], "--selected", kfiles[2],
            "--threshold", 3, "-w", fw2)
    selected = orc.load_kmers([per_sample["meta_test_3"][0]], 0)
    acc = orc.presence_for_kmers([k for _, c in comps for k in c], orc.load_kmers_bin(per_sample["meta_test_1"][0]))
    vec, breadth, _, _ = orc.features(comps, acc, 3, selected)
    assert open(fw2 / "vectors" / "meta_test_1.vec").read() == orc.vec_text(vec)
    assert open(fw2 / "vectors" / "meta_test_1.breadth").read() == orc.breadth_text(breadth)
    acc = orc.presence_for_reads([k for _, c in comps for k in c], orc.parse_reads(files[1]), 31)
    vec, breadth, _, _ = orc.features(comps, acc, 3, selected)
    assert open(fw2 / "vectors" / "meta_test_2.vec").read() == orc.vec_text(vec)
    assert open(fw2 / "vectors" / "meta_test_2.breadth").read() == orc.breadth_text(breadth)


@pytest.mark.parametrize("gpus", [2, 8])
def test_features_calculator_multi_gpu(built, tmp_path, gpus):
    """features-calculator --gpus G on records inputs: the component set on every (logical) GPU, every file's records
    shared out between them, per-component sums added up -- .vec / .breadth byte-equal to the oracle's (= one GPU's)."""
    rng = np.random.default_rng(23)
    files = [os.path.join(INPUTS, "meta_test_%d.fa" % n) for n in (1, 2, 3)]
    per_sample = orc.kmer_counter_many(files, 31, 1)
    kdir = tmp_path / "kmers"
    kdir.mkdir()
    for name, (rec, stat, counts) in per_sample.items():
        (kdir / (name + ".kmers.bin")).write_bytes(rec)
    all_keys = sorted(set().union(*[set(c) for _, _, c in per_sample.values()]))
    comps = [(int(rng.integers(0, 1000)), [all_keys[int(i)] for i in rng.integers(0, len(all_keys), int(rng.integers(1, 600)))]) for _ in range(80)]
    comps.append((0, []))
    cm = tmp_path / "components.bin"
    cm.write_bytes(orc.save_components(comps))
    kfiles = [str(kdir / ("meta_test_%d.kmers.bin" % n)) for n in (1, 2, 3)]
    env = dict(os.environ, MFKC_LOGICAL_GPUS="1")
    for thr, sel in ((0, None), (2, kfiles[1])):
        fw = tmp_path / ("fw%d" % thr)
        cmd = [CLI, "-t", "features-calculator", "-k", "31", "-cm", str(cm), "-ka", *kfiles, "-w", str(fw), "--gpus", str(gpus), "--threshold", str(thr)]
        if sel:
            cmd += ["--selected", sel]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
        assert r.returncode == 0, r.stderr
        selected = orc.load_kmers([per_sample["meta_test_2"][0]], 0) if sel else None
        for n in (1, 2, 3):
            acc = orc.presence_for_kmers([k for _, c in comps for k in c], orc.load_kmers_bin(per_sample["meta_test_%d" % n][0]))
            vec, breadth, _, _ = orc.features(comps, acc, thr, selected)
            assert open(fw / "vectors" / ("meta_test_%d.vec" % n)).read() == orc.vec_text(vec)
            assert open(fw / "vectors" / ("meta_test_%d.breadth" % n)).read() == orc.breadth_text(breadth)


def test_set_algebra_tools_cli(built, tmp_path):
    """kmers-filter, unique-kmers-multi, kmers-samples-counter on the .kmers.bin files the counter wrote:
    same options, default locations, output names and log lines as the reference tools."""
    files = [os.path.join(INPUTS, "meta_test_%d.fa" % n) for n in (1, 2, 3)]
    wd = tmp_path / "wd"
    run_cli("-t", "kmer-counter-many", "-k", 31, "-b", 0, "-i", *files, "-w", wd)
    kf = [str(wd / "kmers" / ("meta_test_%d.kmers.bin" % n)) for n in (1, 2, 3)]
    data = [open(f, "rb").read() for f in kf]
    # kmers-filter: test set = sample 1 and 2, known samples = sample 3
    w1 = tmp_path / "w1"
    r = run_cli("-t", "kmers-filter", "-k", 31, "-i", kf[0], kf[1], "--filter-kmers", kf[2], "-b", 2, "-w", w1)
    want = orc.kmers_filter(data[:2], [data[2]], 2, 0)
    for n, (size, rec) in zip((1, 2), want):
        assert open(w1 / "kmers" / ("meta_test_%d.kmers.bin" % n), "rb").read() == rec and len(rec) > 0
    assert "%s k-mers found" % f"{want[0][0]:,}".replace(",", "'") in r.stderr and "of them survived after filtering" in r.stderr
    # unique-kmers-multi
    w2 = tmp_path / "w2"
    run_cli("-t", "unique-kmers-multi", "-k", 31, "-i", kf[0], kf[1], "--filter-kmers", kf[2], "--min-samples", 1, "--max-samples", 2, "-w", w2)
    size, per_i = orc.unique_kmers_multi(data[:2], [data[2]], 1, 1, 2)
    for i in (1, 2):
        assert open(w2 / "kmers" / ("filtered_%d.kmers.bin" % i), "rb").read() == per_i[i]
    assert len(per_i[1]) > len(per_i[2])
    assert run_cli("-t", "unique-kmers-multi", "-k", 31, "-i", kf[0], "--filter-kmers", kf[2], "--min-samples", 3, "--max-samples", 2,
                   ok=False).returncode == 1
    # kmers-samples-counter
    w3 = tmp_path / "w3"
    r = run_cli("-t", "kmers-samples-counter", "-k", 31, "-i", *kf, "-w", w3)
    size, rec, stat = orc.kmers_samples_counter(data, 1)
    assert open(w3 / "kmers" / "n_samples.kmers.bin", "rb").read() == rec
    assert open(w3 / "stats" / "n_samples.stat.txt").read() == stat
    assert "of them is good (not erroneous)" in r.stderr


def test_min_seq_len_cli(built, tmp_path):
    """the counting call of component-cutter's front half: IOUtils.loadReads(sequences, k, minLen)"""
    fa = tmp_path / "contigs.fa"
    rng = np.random.default_rng(5)
    seqs = ["".join(rng.choice(list("ACGT"), int(n))) for n in rng.integers(20, 400, 80)]
    fa.write_text("".join(">c%d\n%s\n" % (i, s) for i, s in enumerate(seqs)))
    run_cli("-t", "kmer-counter", "-k", 21, "-b", 0, "-l", 100, "-i", fa, "-w", tmp_path / "w")
    want = orc.kmers_bin(orc.count_reads(seqs, 21, 100), 0, 21)
    assert open(tmp_path / "w" / "kmers" / "contigs.kmers.bin", "rb").read() == want
    assert want != orc.kmers_bin(orc.count_reads(seqs, 21, 0), 0, 21)


def test_seq_builder_cli(built, tmp_path):
    """seq-builder on the counter's output: same options, output names and FASTA layout as the reference;
    the sequences (the reference emits them in thread order) come in ascending start-k-mer order."""
    files = [os.path.join(INPUTS, "meta_test_%d.fa" % n) for n in (1, 3)]
    wd = tmp_path / "wd"
    run_cli("-t", "kmer-counter-many", "-k", 31, "-b", 0, "-i", *files, "-w", wd)
    kf = [str(wd / "kmers" / ("meta_test_%d.kmers.bin" % n)) for n in (1, 3)]
    data = [open(f, "rb").read() for f in kf]
    w1 = tmp_path / "w1"
    r = run_cli("-t", "seq-builder", "-k", 31, "-i", kf[0], "-b", 2, "-l", 100, "-w", w1)
    hm = orc.load_kmers(data[:1], 2)
    seqs = orc.seq_builder(hm, 31, 2, 100)
    assert len(seqs) > 3
    assert open(w1 / "sequences" / "meta_test_1.seq.fasta").read() == orc.sequences_fasta(seqs)
    assert open(w1 / "distribution").read() == orc.seq_builder_distribution(hm)
    assert "%d sequences found" % len(seqs) in r.stderr.replace("'", "")
    # two inputs ("+" name), --bottom-cut-percent instead of -b
    w2 = tmp_path / "w2"
    r = run_cli("-t", "seq-builder", "-k", 31, "-i", *kf, "-bp", 10, "--sequence-len", 60, "-o", w2 / "out", "-w", w2)
    hm = orc.load_kmers(data, 1)
    total = sum(hm.values()); stat = [0] * 1024
    for v in hm.values():
        stat[min(v, 1023)] += 1
    b, cur = 1, 0
    for i in range(1023):
        if cur >= total * 10 // 100:
            b = i
            break
        cur += i * stat[i]
    assert "Using maximal bad frequency = %d" % b in r.stderr
    assert open(w2 / "out" / "meta_test_1+.seq.fasta").read() == orc.sequences_fasta(orc.seq_builder(hm, 31, b, 60))


def test_matrix_builder_pipeline_reference_golden(built, tmp_path):
    """The reference's own checked-in result, test_data/meta_test_matrix.txt (= tests/golden/meta_test_matrix.txt): the
    Bray-Curtis matrix of `matrix-builder -k 31 -i meta_test_{1,2,3}.fa` (README.md:90-99).  Every stage that is in
    scope runs on the GPU through mfkc_cli -- kmer-counter-many, seq-builder per sample, component-cutter's counting call
    (kmer-counter -l 100 over the three sequence files), features-calculator -- and only the component split
    (ComponentsBuilder, out of scope) and the distance formula come from the checker.  The three distances must equal
    the reference's to the last bit."""
    from tests.conftest import GOLDEN
    gold = orc.load_matrix_txt(open(os.path.join(GOLDEN, "meta_test_matrix.txt")).read())
    files = [os.path.join(INPUTS, "meta_test_%d.fa" % n) for n in (1, 2, 3)]
    names = ["meta_test_%d" % n for n in (1, 2, 3)]
    wd = tmp_path / "wd"
    run_cli("-t", "kmer-counter-many", "-k", 31, "-i", *files, "-w", wd)                       # default -b 1
    kfiles = [str(wd / "kmers" / (name + ".kmers.bin")) for name in names]
    for kf in kfiles:
        run_cli("-t", "seq-builder", "-k", 31, "-b", 1, "-l", 100, "-i", kf, "-o", wd / "sequences", "-w", wd / "sub-builder")
    sfiles = [str(wd / "sequences" / (name + ".seq.fasta")) for name in names]
    run_cli("-t", "kmer-counter", "-k", 31, "-b", 0, "-l", 100, "-i", *sfiles, "-w", wd / "cutter")
    (seq_kmers,) = os.listdir(wd / "cutter" / "kmers")
    seq_hm = dict(orc.load_kmers_bin(open(wd / "cutter" / "kmers" / seq_kmers, "rb").read()))
    assert len(seq_hm) == 17061
    comps3 = orc.component_cutter(seq_hm, 31, 1000, 10000)
    assert [(len(keys), w, thr) for w, keys, thr in comps3] == \
        [(6240, 12783, 1), (5713, 11265, 1), (3020, 5977, 1), (2088, 4260, 1)]
    cm = wd / "components.bin"
    cm.write_bytes(orc.save_components([(w, keys) for w, keys, _ in comps3]))
    run_cli("-t", "features-calculator", "-k", 31, "-cm", cm, "-ka", *kfiles, "-w", wd)
    vecs = [[float(x) for x in open(wd / "vectors" / (name + ".vec")).read().split()] for name in names]
    assert [int(x) for x in vecs[1]] == [20208, 0, 0, 11337]
    for i in range(3):
        for j in range(3):
            if i != j:
                assert orc.bray_curtis(vecs[i], vecs[j]) == gold[(names[i], names[j])]
    assert orc.java_double_to_string(orc.bray_curtis(vecs[0], vecs[2])) == "0.2981399448537721"


def test_matrix_builder_cli_reproduces_the_reference_file(built, tmp_path):
    """ONE command, the reference's default pipeline (src/tools/DistanceMatrixBuilderMain.java:88-176) on its own test
    samples: the final matrix file equals test_data/meta_test_matrix.txt (tests/golden/) byte for byte.  Counting,
    seq-builder, component-cutter (count + graph split) and features run on the GPU; the distances and the renumbering
    are host arithmetic on 3 x 4 numbers."""
    import glob
    from tests.conftest import GOLDEN
    files = [os.path.join(INPUTS, "meta_test_%d.fa" % n) for n in (2, 3, 1)]
    wd = tmp_path / "wd"
    r = run_cli("-t", "matrix-builder", "-k", 31, "-i", *files, "-w", wd, "--output-format", "%s")
    (out,) = r.stdout.split()
    assert out.startswith(str(wd / "matrices" / "dist_matrix_")) and not out.endswith("_original_order.txt")
    assert open(out).read() == open(os.path.join(GOLDEN, "meta_test_matrix.txt")).read()
    (orig,) = glob.glob(str(wd / "matrices" / "*_original_order.txt"))
    assert open(orig).read().splitlines()[0] == "#\tmeta_test_1\tmeta_test_2\tmeta_test_3"
    # the stages' files, at the reference's default locations
    assert open(wd / "components-stat-1000-10000.txt").read() == (
        "# component.no\tcomponent.size\tcomponent.weight\tusedFreqThreshold\n"
        "1\t6240\t12783\t1\n2\t5713\t11265\t1\n3\t3020\t5977\t1\n4\t2088\t4260\t1\n")
    comps = orc.load_components(open(wd / "components.bin", "rb").read())
    assert [(w, len(keys)) for w, keys in comps] == [(12783, 6240), (11265, 5713), (5977, 3020), (4260, 2088)]
    assert all(keys == sorted(keys) for _, keys in comps)
    assert open(wd / "vectors" / "meta_test_2.vec").read() == "20208\n0\n0\n11337\n"
    assert os.path.exists(wd / "sub-builder" / "distribution") and os.path.exists(wd / "sequences" / "meta_test_3.seq.fasta")
    assert "Total 4 components were found" in r.stderr and "Found 3 libraries to process" in r.stderr
    # default format = the README's table (README.md:96-99); the component-cutter tool on its own with other limits
    r = run_cli("-t", "matrix-builder", "-i", *files, "-w", tmp_path / "w2")
    assert open(r.stdout.split()[0]).read().splitlines()[1] == "meta_test_1\t0.0000\t0.2981\t0.5691"
    sfiles = sorted(glob.glob(str(wd / "sequences" / "*.seq.fasta")))
    run_cli("-t", "component-cutter", "-k", 31, "-i", *sfiles, "-b1", 100, "-b2", 3000, "-l", 100, "-w", tmp_path / "w3",
            "--components-file", tmp_path / "w3" / "c.bin")
    seq_reads = []
    for f in sfiles:
        seq_reads += orc.parse_reads(f)
    want = orc.component_cutter(orc.count_reads(seq_reads, 31, 100), 31, 100, 3000)
    assert len(want) == 34
    assert open(tmp_path / "w3" / "c.bin", "rb").read() == orc.save_components([(w, keys) for w, keys, _ in want])
    assert open(tmp_path / "w3" / "components-stat-100-3000.txt").read() == orc.components_stat_txt(want)
