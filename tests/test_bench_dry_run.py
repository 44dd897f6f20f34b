"""bench.py's own control flow (N = 1) on a machine without a GPU: the device is replaced by tests/_fake_device.py, so
the script's parity checks, warm-up / timed loops, the pipelined end-to-end pass and the JSON line run for real on a
few thousand reads.  Numbers are meaningless here; the GPU run is what `gpurun ... python bench.py` and the driver do."""
import io
import json
import sys
from contextlib import redirect_stdout

import pytest


@pytest.mark.parametrize("serial_e2e", [False, True])
def test_bench_main_runs_through_on_a_fake_device(built, monkeypatch, serial_e2e):
    import bench
    import metafast_b200.sharded                                 # noqa: F401  (resolved before the package is swapped)
    from tests import _fake_device
    fake = _fake_device.module()
    _fake_device.FakeCounter.calls.clear()
    monkeypatch.setitem(sys.modules, "metafast_b200", fake)
    monkeypatch.setattr(bench, "N_READS", 6000)
    monkeypatch.setattr(bench, "BATCH_READS", 1000)
    monkeypatch.setattr(bench, "CPU_SAMPLE_READS", 2000)
    monkeypatch.setattr(bench, "B_THRESHOLD", 0)                 # (6000 reads of a 150-Mbp community: no k-mer is seen three times)
    monkeypatch.setenv("MFKC_BENCH_VERIFY_READS", "3000")
    monkeypatch.setenv("MFKC_BENCH_NO_INGEST", "1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MFKC_BENCH_VARIANT", "MFKC_BENCH_NO_VERIFY", "MFKC_BENCH_NO_CPU", "MFKC_EXCHANGE"):
        monkeypatch.delenv(k, raising=False)
    if serial_e2e:
        monkeypatch.setenv("MFKC_BENCH_E2E_SERIAL", "1")
        monkeypatch.setattr(bench, "load_traffic", lambda: 19.1e9)      # as on the real workload: a committed ncu figure
    else:
        monkeypatch.delenv("MFKC_BENCH_E2E_SERIAL", raising=False)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "4", "--warmup", "3"])
    # cpu_reference_run / verify import the real package's host helpers through `import metafast_b200 as m` too: the fake
    # forwards everything but KmerCounter and load()
    out = io.StringIO()
    with redirect_stdout(out):
        bench.main()
    lines = [ln for ln in out.getvalue().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    # the driver's contract
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["error"] is None and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 32768 * 8
    assert ("one sample after the other" in e["mode"]) == serial_e2e
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["kernel"] == "extract_skm+bin_count" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    if serial_e2e:
        assert r["traffic"] == 19.1e9 and r["traffic_frac_of_peak"] > 0
    else:
        assert r["traffic"] is None and r["traffic_frac_of_peak"] is None     # the committed ncu traffic belongs to the 20 M-read workload only
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0
    v = d["verified"]
    assert v["gpu_equals_cpu_oracle"] is True and v["variants_agree_full_workload"] is True
    assert v["variants_compared"] == ["default", "table", "sort", "direct"]
    assert d["result"]["records_gt_b"] == v["full_workload"]["records"] == v["full_workload"]["distinct"] > 500000
    assert e["d2h_bytes_per_step"] == 10 * d["result"]["records_gt_b"] + 32768 * 8
    # API order per context: nothing is submitted between flush and the next reset, results only after flush
    state = {}
    for cid, what, *rest in _fake_device.FakeCounter.calls:
        if what == "reset":
            state[cid] = "open"
        elif what in ("submit", "submit_device"):
            assert state.get(cid, "open") == "open"
        elif what == "flush":
            state[cid] = "flushed"
        elif what in ("emit_begin", "histogram"):
            assert state.get(cid) == "flushed"


def test_reference_arm_line(built):
    """`bench.py --impl reference` needs no GPU: the C restatement of the reference's CPU algorithm on a bounded sample.
    Under torchrun only rank 0 works and prints."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MFKC_BENCH_CPU_READS="20000")
    env.pop("RANK", None)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 2 and d["value"] > 0 and d["unit"] == "kmers/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["same_config"] is False and 0 < d["config"]["sample_fraction"] < 1
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=600)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


@pytest.mark.parametrize("world", [2, 3])
def test_bench_sharded_runs_through_on_fake_devices(built, tmp_path, world):
    """N > 1: one process per rank, gloo for NCCL, files for peer memory.  Covers the sharded verification, the timed loop
    and the end-to-end pass with several lanes per rank (per-lane process groups, exchange outside the link lock)."""
    import os
    import socket
    import torch.multiprocessing as mp
    from tests import _gloo_worker
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_gloo_worker.run_bench_fake_device, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [open(os.path.join(tmp_path, "bench_rank%d.out" % r)).read() for r in range(world)]
    assert all(o.strip() == "" for o in outs[1:])               # rank 0 alone prints
    lines = [ln for ln in outs[0].splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == world and d["scaling"] == "weak" and d["value"] > 0
    assert d["verified"]["sharded_equals_single_gpu"] is True and d["verified"]["sharded_sample"]["records"] > 100000
    e = d["e2e"]
    assert e["value"] > 0 and e["error"] is None and "3 contexts take the samples in turn" in e["mode"]
    r = d["roofline"]
    assert r["kernel"] == "extract_skm_shard+bin_count" and r["traffic"] is None and 0 < r["frac"]
    assert d["cpu_baseline"] is None
    # every rank's lanes: stage reset -> submissions -> totals -> drain, in that order, per context
    for rank in range(world):
        state = {}
        for ln in open(os.path.join(tmp_path, "calls_rank%d.txt" % rank)).read().splitlines():
            cid, what = ln.split()[:2]
            if what == "p2p_stage_reset":
                state[cid] = "staging"
            elif what in ("p2p_extract", "p2p_submit"):
                assert state.get(cid) == "staging", ln
            elif what == "p2p_counts":
                assert state.get(cid) == "staging", ln
                state[cid] = "counted"
            elif what == "p2p_drain":
                assert state.get(cid) == "counted", ln
                state[cid] = "drained"


def test_cpu_baseline_prefers_a_reference_jvm_when_the_box_has_one(built, tmp_path, monkeypatch):
    """the CPU baseline is the reference JVM if `java` and a complete MetaFast jar exist, else the C
    restatement.  Neither exists in this image: a stand-in `java` (a script that writes the output file) checks the command
    line bench.py would run, and the fall-backs."""
    import os
    import stat
    import bench
    monkeypatch.delenv("JAVA_HOME", raising=False)
    monkeypatch.delenv("METAFAST_JAR", raising=False)
    if not os.path.exists(os.path.join(bench.ROOT, "baseline", "_ref", "metafast.jar")):
        assert bench.jvm_reference_run(1000, 2) is None                      # nothing to run here
    kmers, dt, kind = bench.cpu_baseline_run(3000, 2)
    assert kind == "port" and kmers > 0 and dt > 0
    fake = tmp_path / "bin"
    fake.mkdir()
    log = tmp_path / "java_args.txt"
    (fake / "java").write_text('#!/bin/sh\necho "$@" > %s\nwhile [ $# -gt 0 ]; do if [ "$1" = "-w" ]; then W="$2"; fi; shift; done\n'
                               'mkdir -p "$W/kmers" && : > "$W/kmers/sample.kmers.bin"\n' % log)
    os.chmod(fake / "java", os.stat(fake / "java").st_mode | stat.S_IEXEC)
    jar = tmp_path / "metafast.jar"
    jar.write_bytes(b"PK")
    monkeypatch.setenv("PATH", str(fake) + os.pathsep + os.environ["PATH"])
    monkeypatch.setenv("METAFAST_JAR", str(jar))
    kmers, dt, kind = bench.cpu_baseline_run(3000, 2)
    assert kind == "reference" and 0 < kmers <= 3000 * 120 and dt > 0
    args = log.read_text().split()
    assert args[:2] == ["-jar", str(jar)] and args[2:10] == ["-t", "kmer-counter-many", "-k", "31", "-b", "2", "-p", "2"]
    assert args[10] == "-i" and args[11].endswith("sample.fastq") and args[12] == "-w"
    # a JVM run that fails falls back to the C restatement
    (fake / "java").write_text("#!/bin/sh\necho boom >&2\nexit 3\n")
    kmers, dt, kind = bench.cpu_baseline_run(3000, 2)
    assert kind == "port"
