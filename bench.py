#!/usr/bin/env python
"""bench.py -- canonical 31-mers counted per second (BASELINE.json metric) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

Workload (BASELINE.json configs[1]): synthetic Illumina-like 150 bp reads, 20 M reads of one
sample (community of 64 random genomes, 150 Mbp, log-normal abundances, 0.1 -> 1 % substitution
ramp, 0.1 % N-reads dropped by the parser rule), k = 31, -b 2.  One STEP = one whole pass of the
hot path over that sample: 2-bit pack + canonical k-mer extraction + counting + histogram +
threshold filter + key-sorted big-endian records.

  value  k-mers/s with the parsed reads already resident in HBM (device time, CUDA events).
  e2e    the same step through the C ABI with HOST (pinned) buffers: every step copies the reads
         host->device and the records + histogram device->host inside the timed region.
  N > 1  one process per GPU (torchrun): every rank extracts the k-mers of ITS OWN 20 M reads
         (weak scaling) and stages them per owner shard in its own HBM; every owner drains its
         segments straight out of the peers' memory over NVLink inside the counting kernel
         (CUDA IPC; MFKC_EXCHANGE=nccl selects the NCCL all-to-all + restage flavour) and counts
         only its own hash range (BASELINE.json configs[3] layout).

The JSON line also carries `roofline` (counting kernels vs the measured HBM peak, plus the
random-sector GUPS peak measured in the same run) and `cpu_baseline` (the C restatement of the
reference's CPU algorithm -- oracle/ref_cpu.c -- on a bounded sample, rank 0, N = 1 only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 31
B_THRESHOLD = 2
READ_LEN = 150
N_READS = int(os.environ.get("MFKC_BENCH_READS", 20_000_000))          # per GPU
BATCH_READS = int(os.environ.get("MFKC_BENCH_BATCH", 1_000_000))
CPU_SAMPLE_READS = int(os.environ.get("MFKC_BENCH_CPU_READS", 1_000_000))
ALGO_BYTES_PER_KMER = 64.0 + 0.25 * READ_LEN / (READ_LEN - K + 1)       # 64.3125


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """DRAM bytes (read + write) of the counting kernels per step, from the committed ncu capture of one step
    (profiles/r2_traffic.json, written by tools/ncu_traffic.py: 20 mark + 20 extract_skm launches + bin_count); None if absent
    or if the run is not the configuration the capture was taken on (other read count, other variant, N > 1)."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if N_READS != 20_000_000 or os.environ.get("MFKC_BENCH_VARIANT") or not os.path.exists(p):
        return None
    try:
        return float(json.load(open(p))["per_step_bytes"]["total_counting"])
    except Exception:
        return None


# ------------------------------------------------------------------ clocks under load
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML from a thread of this process (a handful of light
    queries every 50 ms); falls back to an `nvidia-smi -lms` child, whose full queries can stall kernel launches for
    milliseconds on a busy host."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples = []          # (sm_mhz, max_mhz, [reasons])
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._run_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "250"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._run_smi, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _run_nvml(self):
        n = self.nvml
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))
        masks = [(nm, getattr(n, a, None) or getattr(n, b, 0)) for nm, a, b in names]
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                r = get_reasons(self.h)
                self.samples.append((float(sm), float(self.max), [nm for nm, m_ in masks if m_ and (r & m_)]))
            except Exception:
                pass
            time.sleep(0.05)

    def _run_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 7:
                try:
                    self.samples.append((float(f[0]), float(f[1]), [n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
                except ValueError:
                    pass

    def stop(self):
        if self.nvml is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            self.t.join(timeout=1)
        sm = sorted(x[0] for x in self.samples)
        reasons = set(r for x in self.samples for r in x[2])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((x[1] for x in self.samples), default=None),
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


# ------------------------------------------------------------------ reference arm (CPU)
def cpu_reference_run(n_reads, threads, seed_first=0):
    """Time the C restatement of the reference's CPU algorithm (oracle/ref_cpu.c: P worker threads,
    synchronized 32768-read dispatcher with the 2-bit pack inside, striped linear-probing maps,
    iteration-order emit + histogram) on `n_reads` reads of the workload.  Returns (kmers, secs)."""
    import numpy as np
    import metafast_b200 as m
    from tests import _oracle_c
    lib = _oracle_c.load()
    cfg = m.synth_cfg()
    raw = m.synth_reads_host(cfg, seed_first, n_reads)
    keep = ~(raw == ord("N")).any(axis=1)                       # the parser drops reads with N
    bases = np.ascontiguousarray(raw[keep]).reshape(-1)
    nk = int(keep.sum())
    offsets = np.arange(nk + 1, dtype=np.uint64) * np.uint64(READ_LEN)
    t0 = time.perf_counter()
    hm = lib.orc_map_new(threads)
    rc = lib.orc_count_reads(hm, bases.ctypes.data, offsets.ctypes.data, nk, K, 0, threads)
    assert rc == 0
    hist = np.zeros(32768, dtype=np.uint64)
    good = lib.orc_emit(hm, B_THRESHOLD, None, 0, hist.ctypes.data, 0)      # count pass (sizes the buffer)
    out = np.empty(max(good, 1) * 10, dtype=np.uint8)
    hist[:] = 0
    lib.orc_emit(hm, B_THRESHOLD, out.ctypes.data, good, hist.ctypes.data, 0)
    dt = time.perf_counter() - t0
    lib.orc_map_free(hm)
    return nk * (READ_LEN - K + 1), dt


def jvm_reference_run(n_reads, threads):
    """SURVEY.md 8(d), first choice for the CPU baseline: the reference itself.  If a JVM and a complete MetaFast jar are on
    the box (`java` on PATH or $JAVA_HOME; baseline/_ref/metafast.jar or $METAFAST_JAR -- the checked-out reference lacks its
    dependency jar), time `java -jar metafast.jar -t kmer-counter-many -k 31 -b 2 -p P -i <the sample as FASTQ>`.
    Returns (kmers, secs) or None when there is nothing to run (the image of this repository: no JVM)."""
    import shutil
    import tempfile
    java = shutil.which("java") or (os.path.join(os.environ["JAVA_HOME"], "bin", "java") if os.environ.get("JAVA_HOME") else None)
    jar = next((p_ for p_ in (os.environ.get("METAFAST_JAR"), os.path.join(ROOT, "baseline", "_ref", "metafast.jar"))
                if p_ and os.path.exists(p_)), None)
    if not java or not os.path.exists(java) or not jar:
        return None
    import metafast_b200 as m
    cli = os.path.join(ROOT, "metafast_b200", "bin", "mfkc_cli")
    d = tempfile.mkdtemp(prefix="mfkc_jvm_")
    try:
        fq = os.path.join(d, "sample.fastq")
        subprocess.run([cli, "gen-reads", fq, str(n_reads), "0"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        kept = sum(len(o) - 1 for b, o in m.read_file(fq))                  # the parser drops the N-reads, like the reference's
        t0 = time.perf_counter()
        r = subprocess.run([java, "-jar", jar, "-t", "kmer-counter-many", "-k", str(K), "-b", str(B_THRESHOLD), "-p", str(threads),
                            "-i", fq, "-w", os.path.join(d, "work")], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0 or not os.path.exists(os.path.join(d, "work", "kmers", "sample.kmers.bin")):
            sys.stderr.write("bench.py: the reference JVM run failed (%s), using the C restatement\n" % (r.stderr.strip().splitlines() or ["no message"])[-1])
            return None
        return kept * (READ_LEN - K + 1), dt
    finally:
        shutil.rmtree(d, ignore_errors=True)


def cpu_baseline_run(n_reads, threads):
    """-> (kmers, secs, kind): the reference JVM when the box has one ("reference"), else oracle/ref_cpu.c ("port")"""
    try:
        got = jvm_reference_run(n_reads, threads)
    except Exception as ex:                                                   # never let the probe cost the baseline
        sys.stderr.write("bench.py: reference JVM probe failed: %r\n" % (ex,))
        got = None
    if got:
        return got[0], got[1], "reference"
    kmers, dt = cpu_reference_run(n_reads, threads)
    return kmers, dt, "port"


def host_ingest_numbers(n_reads=400_000):
    """Throughput of the host ingest path (mfkc_reader_*: mapped input, parallel parsers, the repository's own multi-threaded
    gzip decoder) on a synthetic FASTQ of the workload's reads, plain and .gz, with zlib's gzread beside it.  Informational:
    real FASTQ(.gz)-to-.kmers.bin runs are bound by this, not by the GPU.  A few seconds, rank 0 at N = 1 only."""
    import gzip
    import shutil
    import tempfile
    import metafast_b200 as m
    cli = os.path.join(ROOT, "metafast_b200", "bin", "mfkc_cli")
    d = tempfile.mkdtemp(prefix="mfkc_ingest_")
    try:
        fq = os.path.join(d, "reads.fastq")
        subprocess.run([cli, "gen-reads", fq, str(n_reads), "0"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        with open(fq, "rb") as f, gzip.open(fq + ".gz", "wb", compresslevel=1) as g:
            shutil.copyfileobj(f, g, 1 << 24)

        import numpy as np
        lib = m.load()
        cap_reads = 1 << 20
        bases = np.ones(128 << 20, dtype=np.uint8)
        offs = np.zeros(cap_reads + 1, dtype=np.uint64)
        got = C.c_uint32()
        err = C.create_string_buffer(256)

        def rate(path, env):
            best = None
            for _ in range(3):
                old = {k: os.environ.get(k) for k in env}
                os.environ.update(env)
                try:
                    h = C.c_void_p()
                    if lib.mfkc_reader_open(path.encode(), C.byref(h), err, 256) != 0:
                        raise RuntimeError(err.value.decode())
                    t0 = time.perf_counter()
                    n = 0
                    while True:                                       # the loop of mfkc_cli: the same buffers over and over
                        if lib.mfkc_reader_next(h, bases.ctypes.data_as(C.c_void_p), bases.nbytes, offs.ctypes.data_as(C.c_void_p),
                                                cap_reads, C.byref(got)) != 0:
                            raise RuntimeError(lib.mfkc_reader_error(h).decode())
                        if got.value == 0:
                            break
                        n += got.value
                    dt = time.perf_counter() - t0
                    lib.mfkc_reader_close(h)
                finally:
                    for k, v in old.items():
                        if v is None:
                            os.environ.pop(k, None)
                        else:
                            os.environ[k] = v
                best = dt if best is None else min(best, dt)
            return n / best
        return {"unit": "reads/s", "cores": os.cpu_count(), "plain_fastq": rate(fq, {}), "fastq_gz": rate(fq + ".gz", {}),
                "fastq_gz_zlib_gzread": rate(fq + ".gz", {"MFKC_INFLATE": "zlib"}),
                "sample": "%d synthetic 150-bp reads as FASTQ (%d MB), gzip -1 (%d MB); best of 3 passes through mfkc_reader_*"
                          % (n_reads, os.path.getsize(fq) >> 20, os.path.getsize(fq + ".gz") >> 20)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import __graft_entry__ as g
    g.build()
    threads = os.cpu_count() or 1
    sample = CPU_SAMPLE_READS
    for _ in range(args.warmup and 1):
        cpu_reference_run(min(sample, 100_000), threads)
    kmers_tot, t_tot, kind = 0, 0.0, "port"
    for s in range(args.steps):
        kmers, dt, kind = cpu_baseline_run(sample, threads)
        kmers_tot += kmers; t_tot += dt
    value = kmers_tot / t_tot
    line = {
        "impl": "reference", "metric": "canonical 31-mers counted/s", "value": value, "unit": "kmers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": dict(workload_config(args.gpus), sample_fraction=sample / N_READS, same_config=False),
        "cpu_baseline": {"value": value, "unit": "kmers/s", "cores": threads, "kind": kind,
                         "sample": ("%d of the %d reads of the workload per step (the reference JVM: kmer-counter-many -p %d on the "
                                    "sample as FASTQ, file parsing included)" % (sample, N_READS, threads)) if kind == "reference" else
                                   ("%d of the %d reads of the workload per step (oracle/ref_cpu.c: C restatement of the "
                                    "reference's threaded algorithm; no JVM in the image)" % (sample, N_READS))},
        "e2e": {"value": value, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "configs[1]: synthetic Illumina 150bp reads, %d reads per GPU (single sample), k=31, -b 2" % N_READS,
            "k": K, "b": B_THRESHOLD, "read_len": READ_LEN, "reads_per_gpu": N_READS, "batch_reads": BATCH_READS,
            "variant": os.environ.get("MFKC_BENCH_VARIANT", "hash (bin-local)"),
            "parallelism": "1 GPU" if n_gpus == 1 else "hash-range sharded over %d GPUs, %s" % (n_gpus, "NCCL all-to-all + restage" if os.environ.get("MFKC_EXCHANGE") == "nccl" else "records drained straight from peer HBM over NVLink (CUDA IPC), no data-path collective"),
            "l2": "inputs (3 GB reads, multi-GB table) are far larger than the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------ parity at size
def verify_at_size(m, kc, step_device, env):
    """Checks, once per run and outside every timed region, that what is timed is right:
      N = 1  (i)  the first MFKC_BENCH_VERIFY_READS (2 M) reads of the workload, counted on the GPU through the host-fed
                  C-ABI path, equal oracle/ref_cpu.c's result: sha256 of the records, histogram, distinct count;
             (ii) the full workload (20 M reads): the default variant, the region-blocked table, the sort-and-RLE variant
                  and the direct-upsert variant produce identical records (sha256), histograms and distinct counts.
      N > 1  (iii) the sharded path as timed (real CUDA IPC between the ranks' processes): the merged shard records and
                  the summed histograms of G x 1 M reads equal ONE single-GPU count of the same reads (rank 0)."""
    import hashlib
    import numpy as np
    world, rank, dist = env["world"], env["rank"], env["dist"]
    out = {}

    def digest(c):
        c.flush()
        rec = c.emit(B_THRESHOLD)
        return hashlib.sha256(rec).hexdigest(), c.histogram(), c.stats()["distinct"], len(rec) // 10

    if world == 1:
        from tests import _oracle_c
        import __graft_entry__ as g
        g.build()
        v_reads = int(os.environ.get("MFKC_BENCH_VERIFY_READS", 2_000_000))
        raw = m.synth_reads_host(m.synth_cfg(), 0, v_reads)
        keep = ~(raw == ord("N")).any(axis=1)
        bases = np.ascontiguousarray(raw[keep]).reshape(-1)
        nk = int(keep.sum())
        offsets = np.arange(nk + 1, dtype=np.uint64) * np.uint64(READ_LEN)
        del raw
        want_rec, want_hist, want_distinct, _ = _oracle_c.count(bases, offsets, K, B_THRESHOLD, P=os.cpu_count() or 1)
        with m.KmerCounter(K, device=env["local_rank"], variant=env["variant"], expected_kmers=nk * (READ_LEN - K + 1)) as kv:
            step = 250_000
            for s in range(0, nk, step):
                kv.submit(bases, offsets[s:min(nk, s + step) + 1])          # host buffers, asynchronous pipeline
            sha, hist, distinct, n_rec = digest(kv)
        out["gpu_equals_cpu_oracle"] = bool(sha == hashlib.sha256(want_rec).hexdigest() and (hist == want_hist).all()
                                            and distinct == want_distinct)
        out["cpu_oracle_sample"] = {"reads": nk, "distinct": int(want_distinct), "records": len(want_rec) // 10}
        del bases, offsets, want_rec
        # (ii) every variant on the full workload
        step_device()
        ref = digest(kc)
        names, agree = ["default"], True
        for name, v in (("table", m.VARIANT_HASH_TABLE), ("sort", m.VARIANT_SORT), ("direct", m.VARIANT_HASH_DIRECT)):
            if v == env["variant"] or name in os.environ.get("MFKC_BENCH_VERIFY_SKIP", "").split(","):
                continue
            with m.KmerCounter(K, device=env["local_rank"], variant=v, expected_kmers=env["kmers_ub"]) as kv:
                n = env["n_reads"]
                for s in range(0, n, BATCH_READS):
                    e = min(n, s + BATCH_READS)
                    kv.submit_device(env["d_bases"] + s * READ_LEN, env["d_offs"] + s * 8, e - s, (e - s) * READ_LEN)
                got = digest(kv)
            names.append(name)
            agree = agree and got[0] == ref[0] and bool((got[1] == ref[1]).all()) and got[2] == ref[2]
        out["variants_agree_full_workload"] = bool(agree)
        out["variants_compared"] = names
        out["full_workload"] = {"sha256_records": ref[0], "records": ref[3], "distinct": int(ref[2])}
        return out

    # (iii) N > 1
    from metafast_b200.sharded import merge_sorted_records
    vr = min(int(os.environ.get("MFKC_BENCH_VERIFY_SHARD_READS", 1_000_000)), env["n_reads"])
    kc.reset()
    if env["exchange"] == "p2p":
        env["sharded"].begin()
    env["sharded"].run_device(env["d_bases"], env["d_offs"], vr)
    kc.flush()
    rec = kc.emit(B_THRESHOLD)
    mine = (rec, kc.histogram(), kc.stats()["distinct"])
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    ok = None
    if rank == 0:
        merged = merge_sorted_records([p[0] for p in parts])
        hist = sum(p[1] for p in parts)
        distinct = sum(int(p[2]) for p in parts)
        with m.KmerCounter(K, device=env["local_rank"], expected_kmers=world * vr * (READ_LEN - K + 1)) as kv:
            span = vr + vr // 50 + 1000
            d_b = kv.device_alloc(span * READ_LEN)
            d_o = kv.device_alloc((span + 1) * 8)
            for r in range(world):                                     # the same reads every rank took: its first vr kept ones
                cfg_r = m.synth_cfg(sample=r)
                kept = C.c_uint64()
                kv._ck(kv.lib.mfkc_synth_reads_device(kv.h, C.byref(cfg_r), r * N_READS, span, C.c_void_p(d_b), C.c_void_p(d_o), C.byref(kept)))
                assert kept.value >= vr
                kv.submit_device(d_b, d_o, vr, vr * READ_LEN)
                kv.sync()
            kv.flush()
            one = kv.emit(B_THRESHOLD)
            ok = bool(one == merged and (kv.histogram() == hist).all() and kv.stats()["distinct"] == distinct)
            kv.device_free(d_b); kv.device_free(d_o)
        out["sharded_equals_single_gpu"] = ok
        out["sharded_sample"] = {"reads_per_gpu": vr, "records": len(merged) // 10, "distinct": distinct,
                                 "sha256_records": hashlib.sha256(merged).hexdigest()}
    return out


# ------------------------------------------------------------------ the other BASELINE configs (not the driver's line)
def _dist_setup():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local_rank = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return rank, world, local_rank, dist


def _max_over_ranks(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    return float(t.item())


def _barrier(dist, kc):
    kc.sync()
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()


def run_config3(args):
    """BASELINE configs[2]: kmer-counter-many batch mode, S samples x R reads (100 x 5 M), k = 31, default -b 1, then
    features-calculator of every sample's records against fixed components (10 000 x 2 000 k-mers).  The samples are
    independent: rank r takes samples r, r + N, ... (no exchange).  value: reads of the rank's samples resident in HBM, the
    records go from the counter to the features tables on the device; e2e: the same through host buffers (reads H2D,
    records D2H, records H2D again as features-calculator's -ka input, vectors D2H), one sample after the other."""
    import numpy as np
    import metafast_b200 as m
    rank, world, local_rank, dist = _dist_setup()
    S = int(os.environ.get("MFKC_BENCH_SAMPLES", 100)); R = int(os.environ.get("MFKC_BENCH_SAMPLE_READS", 5_000_000))
    n_comp, comp_size, b = 10_000, 2_000, 1
    mine = list(range(rank, S, world))
    kc = m.KmerCounter(K, device=local_rank, expected_kmers=R * (READ_LEN - K + 1))
    fc = m.FeaturesCalculator(K, device=local_rank)
    bufs = []
    for sidx in mine:                                               # every sample: its own abundance vector of the community
        d_b = kc.device_alloc(R * READ_LEN); d_o = kc.device_alloc((R + 1) * 8)
        kept = C.c_uint64(); cfg = m.synth_cfg(sample=sidx)
        kc._ck(kc.lib.mfkc_synth_reads_device(kc.h, C.byref(cfg), 0, R, C.c_void_p(d_b), C.c_void_p(d_o), C.byref(kept)))
        bufs.append((d_b, d_o, kept.value))

    def count(d_b, d_o, n):
        kc.reset()
        for s0 in range(0, n, BATCH_READS):
            e = min(n, s0 + BATCH_READS)
            kc.submit_device(d_b + s0 * READ_LEN, d_o + s0 * 8, e - s0, (e - s0) * READ_LEN)
        kc.flush()
        return kc.emit_begin(b)

    # fixed components: consecutive runs of sample 0's records (the same on every rank)
    kc0 = kc
    d_b0 = kc0.device_alloc(R * READ_LEN); d_o0 = kc0.device_alloc((R + 1) * 8)
    kept = C.c_uint64(); cfg0 = m.synth_cfg(sample=0)
    kc0._ck(kc0.lib.mfkc_synth_reads_device(kc0.h, C.byref(cfg0), 0, R, C.c_void_p(d_b0), C.c_void_p(d_o0), C.byref(kept)))
    count(d_b0, d_o0, kept.value)
    rec0 = np.frombuffer(kc.emit(b), dtype=np.uint8).reshape(-1, 10)
    kc0.device_free(d_b0); kc0.device_free(d_o0)
    n_keys = min(len(rec0), n_comp * comp_size)
    comp_size = max(1, n_keys // n_comp)
    keys = np.ascontiguousarray(rec0[: n_comp * comp_size, :8]).view(">u8").astype(np.uint64).reshape(-1)
    off = (np.arange(n_comp + 1, dtype=np.uint64) * np.uint64(comp_size))
    fc._ck(fc.lib.mfkc_fc_load_components(fc.h, keys.view(np.int64).ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), n_comp))
    fc.n_comp = n_comp
    del rec0

    kmers_mine = sum(n * (READ_LEN - K + 1) for _, _, n in bufs)
    recs_mine = [0]

    def step():
        recs_mine[0] = 0
        for d_b, d_o, n in bufs:
            recs_mine[0] += count(d_b, d_o, n)
            fc.reset_values()
            fc.add_emitted(kc)
            vec, found, cnt = fc.features(0)
        return vec

    for _ in range(max(1, min(args.warmup, 1))):
        step()
    _barrier(dist, kc)
    kc.profile(enable=True, reset=True); fc.profile(enable=True, reset=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vec = step()
    _barrier(dist, kc)
    dt = _max_over_ranks(dist, time.perf_counter() - t0)
    prof = kc.profile(enable=False); fprof = fc.profile(enable=False)
    kmers_all = _sum_over_ranks(dist, kmers_mine)
    # e2e through host buffers, the rank's first two samples in turn
    import numpy as _np
    host = []
    for d_b, d_o, n in bufs[:2]:
        hb = kc.pinned(n * READ_LEN); ho = kc.pinned((n + 1) * 8, _np.uint64)
        kc.d2h(hb, d_b); kc.d2h(ho, d_o); host.append((hb, ho, n))
    h_out = kc.pinned(max(int(recs_mine[0] / max(1, len(bufs)) * 1.3) + 4096, 4096) * 10)

    def e2e_sample(hb, ho, n):
        kc.reset()
        for s0 in range(0, n, BATCH_READS):
            kc.submit(hb, ho[s0:min(n, s0 + BATCH_READS) + 1])
        kc.flush()
        nbytes = kc.emit_into(b, h_out)
        kc.histogram()
        fc.reset_values()
        chunk = 16777200                                              # the reference's KMERS_WORK_RANGE_SIZE (src/io/IOUtils.java:30)
        for pos in range(0, nbytes, chunk):
            nb = min(chunk, nbytes - pos)
            fc._ck(fc.lib.mfkc_fc_add_records(fc.h, C.c_void_p(h_out.ctypes.data + pos), nb // 10))
        fc.features(0)
        return nbytes
    for hb, ho, n in host:
        e2e_sample(hb, ho, n)
    _barrier(dist, kc)
    t0 = time.perf_counter()
    reps = max(1, len(bufs))
    for i in range(reps):
        out_bytes = e2e_sample(*host[i % len(host)])
    _barrier(dist, kc)
    e2e_dt = _max_over_ranks(dist, time.perf_counter() - t0)
    if rank == 0:
        peak, peak_src = load_peaks()
        fc_ms = fprof.get("fc_records", (0, 0))[0] / args.steps
        hits = float(n_comp * comp_size)                          # upper bound: every component k-mer met once per sample
        recs = recs_mine[0]
        fc_bytes = 10.0 * recs + 64.0 * hits * len(bufs)
        line = {"metric": "canonical 31-mers counted/s", "value": kmers_all * args.steps / dt, "unit": "kmers/s", "n_gpus": world, "steps": args.steps,
                "warmup": 1, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "int64", "data": "synthetic",
                "config": {"workload": "configs[2]: %d synthetic samples x %d reads, k=31, -b 1, then features-calculator on %d components x %d k-mers"
                                       % (S, R, n_comp, comp_size), "parallelism": "samples round robin over %d GPU(s), no exchange" % world},
                "samples_per_s": S * args.steps / dt, "ms_per_sample": 1e3 * dt / args.steps / max(1, len(mine)),
                "e2e": {"value": kmers_mine / max(1, len(bufs)) * reps * world / e2e_dt, "unit": "kmers/s", "ms_per_sample": 1e3 * e2e_dt / reps,
                        "h2d_bytes_per_step": int(host[0][2] * READ_LEN + out_bytes), "d2h_bytes_per_step": int(out_bytes + 32768 * 8 + 3 * 8 * n_comp),
                        "mode": "one sample after the other per GPU"},
                "kernel_ms_per_step": {k_: v[0] / args.steps for k_, v in list(prof.items()) + list(fprof.items()) if v[1] and v[0]},
                "features_roofline": {"bound": "hbm", "kernel": "fc_pairs", "achieved": fc_bytes / (fc_ms / 1e3) / 1e9 if fc_ms else None, "peak": peak, "unit": "GB/s",
                                      "frac": (fc_bytes / (fc_ms / 1e3) / 1e9 / peak) if fc_ms else None,
                                      "note": "10 B per record streamed + 64 B per hit (SURVEY 8d); records %d per rank and step" % recs},
                "gpu_launches": int(sum(v[1] for v in prof.values()) + sum(v[1] for v in fprof.values()))}
        print(json.dumps(line))
    kc.close(); fc.close()
    if dist is not None:
        dist.destroy_process_group()


def run_config_sharded(args, k, config_name, total_reads, paired=False):
    """BASELINE configs[3] (k = 31, one metagenome hash-range sharded over the GPUs) and configs[4] (k = 55, 128-bit keys,
    paired-end files): ONE sample of `total_reads` reads, split between the ranks (strong scaling); every rank extracts its
    slice, every owner counts its hash range out of the peers' staging buffers, filter + histogram per shard."""
    import numpy as np
    import metafast_b200 as m
    rank, world, local_rank, dist = _dist_setup()
    per_rank = total_reads // world
    kmers_ub = per_rank * (READ_LEN - k + 1)
    kc = m.KmerCounter(k, device=local_rank, expected_kmers=kmers_ub, n_shards=world if world > 1 else 0, shard_id=rank if world > 1 else 0)
    cfg = m.synth_cfg(sample=0)
    d_b = kc.device_alloc(per_rank * READ_LEN); d_o = kc.device_alloc((per_rank + 1) * 8)
    kept = C.c_uint64()
    # paired-end: mates are consecutive reads of the generator (R1 = even, R2 = odd); a rank takes whole pairs
    kc._ck(kc.lib.mfkc_synth_reads_device(kc.h, C.byref(cfg), rank * per_rank, per_rank, C.c_void_p(d_b), C.c_void_p(d_o), C.byref(kept)))
    n = kept.value
    sharded = None
    if world > 1:
        from metafast_b200.sharded import P2PShardedStep
        sharded = P2PShardedStep(kc, dist, world, rank, BATCH_READS, READ_LEN, k, per_rank)

    def step():
        kc.reset()
        if sharded:
            sharded.begin()
            sharded.run_device(d_b, d_o, n)
        else:
            for s0 in range(0, n, BATCH_READS):
                e = min(n, s0 + BATCH_READS)
                kc.submit_device(d_b + s0 * READ_LEN, d_o + s0 * 8, e - s0, (e - s0) * READ_LEN)
        kc.flush()
        return kc.emit_begin(B_THRESHOLD)

    for _ in range(args.warmup):
        good = step()
    _barrier(dist, kc)
    kc.profile(enable=True, reset=True)
    kc.timer_start()
    for _ in range(args.steps):
        good = step()
    ms = _max_over_ranks(dist, kc.timer_stop_ms())
    _barrier(dist, kc)
    prof = kc.profile(enable=False)
    st = kc.stats()
    kmers_all = _sum_over_ranks(dist, n * (READ_LEN - k + 1))
    distinct_all = _sum_over_ranks(dist, st["distinct"]); good_all = _sum_over_ranks(dist, good)
    if rank == 0:
        line = {"metric": "canonical %d-mers counted/s" % k, "value": kmers_all * args.steps / (ms / 1e3), "unit": "kmers/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "int64" if k <= 31 else "int128", "data": "synthetic",
                "config": {"workload": "%s: ONE sample of %d synthetic 150-bp reads%s, k=%d, -b %d" % (config_name, total_reads, " (paired-end, _R1/_R2)" if paired else "", k, B_THRESHOLD),
                           "reads_per_gpu": per_rank, "parallelism": "hash-range sharded over %d GPU(s); %s" % (world, "bin-local count out of peer memory" if k <= 31 else "region-blocked 128-bit table drained from peer memory")},
                "kernel_ms_per_step": {k_: v[0] / args.steps for k_, v in prof.items() if v[1] and v[0]},
                "gpu_launches": int(sum(v[1] for v in prof.values())),
                "result": {"kmers": kmers_all, "distinct": distinct_all, "records_gt_b": good_all, "bins": kc.bin_stats()}}
        print(json.dumps(line))
    kc.close()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------ this repository's arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mfkc", choices=["mfkc", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json configs index + 1: 2 = the metric's configuration (default, the driver's line); 3 = batch of "
                         "samples + features; 4 = one metagenome sharded over the GPUs (strong scaling); 5 = k = 55 paired-end")
    ap.add_argument("--reads", type=int, default=0, help="configs 4 / 5: total reads of the one sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.config == 3:
        return run_config3(args)
    if args.config == 4:      # 2 G reads in BASELINE; default here = 10x scaled down (what fits device-resident at N >= 2)
        return run_config_sharded(args, 31, "configs[3]", args.reads or 200_000_000)
    if args.config == 5:      # 500 M reads in BASELINE; default here = 10x scaled down
        return run_config_sharded(args, 55, "configs[4]", args.reads or 50_000_000, paired=True)

    import numpy as np
    import metafast_b200 as m

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if m.load().mfkc_device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device (libmfkc has no CPU fallback)")

    variant = {"sort": m.VARIANT_SORT, "direct": m.VARIANT_HASH_DIRECT, "table": m.VARIANT_HASH_TABLE}.get(os.environ.get("MFKC_BENCH_VARIANT", ""), m.VARIANT_HASH)
    kmers_ub = N_READS * (READ_LEN - K + 1)
    kc = m.KmerCounter(K, device=local_rank, variant=variant, expected_kmers=kmers_ub,
                       n_shards=world if world > 1 else 0, shard_id=rank if world > 1 else 0)
    cfg = m.synth_cfg(sample=rank)                                   # every rank: its own reads of the community
    # ---- synthetic parsed reads, device-resident
    d_bases = kc.device_alloc(N_READS * READ_LEN)
    d_offs = kc.device_alloc((N_READS + 1) * 8)
    kept = C.c_uint64()
    kc._ck(kc.lib.mfkc_synth_reads_device(kc.h, C.byref(cfg), rank * N_READS, N_READS, C.c_void_p(d_bases),
                                          C.c_void_p(d_offs), C.byref(kept)))
    n_reads = kept.value
    n_bases = n_reads * READ_LEN
    kmers_per_step = n_reads * (READ_LEN - K + 1)

    exchange = os.environ.get("MFKC_EXCHANGE", "p2p")                # "p2p": peer-memory drain (default); "nccl": all-to-all + restage
    if world > 1:
        from metafast_b200.sharded import ShardedStep, P2PShardedStep
        if exchange == "p2p":
            sharded = P2PShardedStep(kc, dist, world, rank, BATCH_READS, READ_LEN, K, N_READS)
        else:
            sharded = ShardedStep(kc, dist, world, rank, BATCH_READS, READ_LEN, K)

    phases = {"reset": 0.0, "submit": 0.0, "flush": 0.0, "emit_begin": 0.0}       # host wall time per phase (where a slow host shows)

    def step_device():
        """one whole pass, inputs resident in HBM"""
        t0 = time.perf_counter()
        kc.reset()
        t1 = time.perf_counter()
        if world > 1:
            if exchange == "p2p":
                sharded.begin()
            sharded.run_device(d_bases, d_offs, n_reads)
        else:
            for s in range(0, n_reads, BATCH_READS):
                e = min(n_reads, s + BATCH_READS)
                kc.submit_device(d_bases + s * READ_LEN, d_offs + s * 8, e - s, (e - s) * READ_LEN)
        t2 = time.perf_counter()
        kc.flush()
        t3 = time.perf_counter()
        n = kc.emit_begin(B_THRESHOLD)
        t4 = time.perf_counter()
        phases["reset"] += t1 - t0; phases["submit"] += t2 - t1; phases["flush"] += t3 - t2; phases["emit_begin"] += t4 - t3
        return n

    def barrier():
        kc.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- parity at size, before anything is timed (the oracle is the checker here, never the thing measured)
    verified = None
    if not os.environ.get("MFKC_BENCH_NO_VERIFY"):
        try:
            verified = verify_at_size(m, kc, step_device, dict(world=world, rank=rank, local_rank=local_rank, dist=dist, variant=variant,
                                                              d_bases=d_bases, d_offs=d_offs, n_reads=n_reads, kmers_ub=kmers_ub,
                                                              sharded=sharded if world > 1 else None, exchange=exchange))
        except Exception as ex:              # the checker broke (not: the check failed): say so in the line, still measure
            verified = {"checker_error": repr(ex)[:300]}
            sys.stderr.write("bench.py: the verification could not run: %r\n" % (ex,))
        if rank == 0 and not all(v for k_, v in verified.items() if isinstance(v, bool)):
            sys.stderr.write("bench.py: VERIFICATION FAILED: %s\n" % json.dumps(verified))

    for _ in range(args.warmup):
        n_good = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    kc.profile(enable=True, reset=True)
    for k_ in phases:
        phases[k_] = 0.0
    kc.timer_start()
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        n_good = step_device()
    ms_total = kc.timer_stop_ms()
    barrier()
    wall = time.perf_counter() - t_wall
    prof = kc.profile(enable=False)
    clocks = sampler.stop()
    ms_total = max_over_ranks(ms_total)
    st = kc.stats()
    value = kmers_per_step * world * args.steps / (ms_total / 1e3)

    # ---- end to end through the C ABI with host buffers (reads H2D, records + histogram D2H every step)
    h_bases = kc.pinned(n_bases)
    h_offs = kc.pinned((n_reads + 1) * 8, np.uint64)
    kc.d2h(h_bases, d_bases)
    kc.d2h(h_offs, d_offs)
    h_out = kc.pinned(max(int(n_good * 1.05) + 1024, 1024) * 10)

    def step_e2e(c=None, out=None, lock=None, shd=None):
        c = c or kc
        out = h_out if out is None else out
        shd = shd or (sharded if world > 1 else None)
        c.reset()
        if world > 1 and exchange == "p2p":
            shd.begin()                      # a barrier between the ranks: never under the link lock
        if lock:
            lock.acquire()                   # one sample at a time on the host-to-device link
        try:
            if world > 1 and exchange == "p2p":
                shd.submit_host(h_bases, h_offs, n_reads)      # this rank's copies + launches only: waits for no other rank
            elif world > 1:
                shd.run_host(h_bases, h_offs, n_reads)         # (collectives inside: only ever run without lanes)
            else:
                for s in range(0, n_reads, BATCH_READS):
                    e = min(n_reads, s + BATCH_READS)
                    c.submit(h_bases, h_offs[s:e + 1])
        finally:
            if lock:
                lock.release()
        if world > 1 and exchange == "p2p":
            # The exchange of the totals waits for the same lane of every other rank.  It must not run under the link lock:
            # rank A's lane 0 holding A's lock in this exchange while rank B's lane 1 holds B's lock in its own is a deadlock
            # (A's lane 1 and B's lane 0 are both queued behind those locks) -- seen as a hung 8-rank run.
            shd.finish()
        c.flush()
        nbytes = c.emit_into(B_THRESHOLD, out)
        c.histogram()
        return nbytes

    # N = 1: the samples of a kmer-counter-many run are independent, so a few contexts take them in turn (what mfkc_cli does
    # in batch mode): while one sample is counted, sorted and copied back, the next one's reads are already crossing PCIe.
    # Every step still does the whole job through the C ABI: reads host -> device, records + histogram device -> host.
    # N > 1: the same with one sharded group of contexts per lane; the lanes' tiny exchanges (G numbers, barriers) run on CPU
    # process groups of their own, so that lanes of different ranks may be at different points.  Those groups time out after
    # a few minutes: a protocol error ends the run with a message instead of hanging it.
    pipelined = not os.environ.get("MFKC_BENCH_E2E_SERIAL") and (world == 1 or (exchange == "p2p" and variant == m.VARIANT_HASH))
    e2e_error = None
    if pipelined:
        import datetime
        from metafast_b200.sharded import run_lanes
        n_lanes = max(2, int(os.environ.get("MFKC_BENCH_E2E_LANES", 3)))
        lane_timeout = datetime.timedelta(seconds=int(os.environ.get("MFKC_BENCH_LANE_TIMEOUT_S", 120)))
        extra = [m.KmerCounter(K, device=local_rank, variant=variant, expected_kmers=kmers_ub,
                               n_shards=world if world > 1 else 0, shard_id=rank if world > 1 else 0) for _ in range(n_lanes - 1)]
        lanes = [(kc, h_out, sharded if world > 1 else None)]
        if world > 1:
            sharded.group = dist.new_group(backend="gloo", timeout=lane_timeout)
        for c in extra:
            shd = None
            if world > 1:
                shd = P2PShardedStep(c, dist, world, rank, BATCH_READS, READ_LEN, K, N_READS,
                                     group=dist.new_group(backend="gloo", timeout=lane_timeout))
            lanes.append((c, c.pinned(h_out.nbytes), shd))
        link = threading.Lock()

        def run_pipelined(n_steps):
            return run_lanes(n_steps, lanes, lambda ln, i: step_e2e(ln[0], ln[1], link, ln[2]))[-1]

        out_bytes = 0
        try:
            run_pipelined(2 * n_lanes)       # every context: plan from a first sample, allocate, warm up
        except RuntimeError as e:            # (a lane group timed out or a context failed: every rank gets here)
            e2e_error = str(e)[:300]
        barrier()
        t0 = time.perf_counter()
        try:
            if e2e_error is None:
                out_bytes = run_pipelined(args.steps)
        except RuntimeError as e:
            e2e_error = str(e)[:300]
        for c in extra:
            c.sync()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0 if e2e_error is None else float("inf"))
        barrier()
        for c in extra:
            c.close()
        barrier()
        if world > 1:
            sharded.group = None
    else:
        for _ in range(2):                       # the first host-fed samples size the staging / table of the context
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_bytes = step_e2e()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_failed = e2e_s == float("inf")                # on any rank (the maximum over the ranks)
    e2e_value = None if e2e_failed else kmers_per_step * world * args.steps / e2e_s

    # what the host can feed: every rank copies its reads host -> device at the same time, nothing else running
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        kc.h2d(d_bases, h_bases)
    barrier()
    link_gbs = 3 * n_bases * world / max_over_ranks(time.perf_counter() - t0) / 1e9
    if rank != 0:
        kc.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the counting kernels (CUDA-event durations recorded around every launch)
    peak, peak_src = load_peaks()
    count_kernels = {"hash (bin-local)": ["extract_skm", "bin_count", "drain_heavy"], "table": ["extract_skm", "drain_skm"],
                     "direct": ["extract_direct"], "sort": ["extract_keys", "radix_sort", "rle"]}[workload_config(1)["variant"]]
    if world > 1:
        # p2p: extract_skm_shard stages by (owner, bin); bin_count (+ drain_heavy) reads the peers' staging over NVLink.
        # drain_skm / extract_skm: the table variant of the exchange (k > 31, MFKC_VARIANT_HASH_TABLE)
        count_kernels = ["extract_skm_shard", "bin_count", "drain_heavy", "extract_skm", "drain_skm"]
    count_kernels = [k for k in count_kernels if k in prof and prof[k][1]]     # those that ran
    t_count_ms = sum(prof[k][0] for k in count_kernels) / args.steps
    launches = sum(v[1] for v in prof.values())
    achieved = ALGO_BYTES_PER_KMER * kmers_per_step / (t_count_ms / 1e3) / 1e9 if t_count_ms else None
    gups_ms = kc.gups(8 << 30, 1 << 28, 1)
    gups_rate = (1 << 28) / (gups_ms / 1e3)
    traffic = load_traffic() if world == 1 else None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
        "traffic": traffic, "peak_source": peak_src,
        # what the DRAM really does: ncu bytes of the counting kernels / their time, as a fraction of the same peak
        "traffic_frac_of_peak": traffic / (t_count_ms / 1e3) / 1e9 / peak if traffic and t_count_ms else None,
        "kernel": "+".join(count_kernels), "algorithmic_bytes_per_kmer": ALGO_BYTES_PER_KMER,
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items() if v[1]},
        "kernel_launches_per_step": {k: v[1] / args.steps for k, v in prof.items() if v[1]},
        "gups_random_sector_upserts_per_s": gups_rate,
        "gups_algorithmic_gbs": gups_rate * 64 / 1e9,
        "frac_of_gups": (kmers_per_step / (t_count_ms / 1e3)) / gups_rate if t_count_ms else None,
        "note": "achieved = 64.3125 B per k-mer instance (SURVEY 8d: one random 32 B sector read + written per instance) / summed "
                "CUDA-event time of the counting kernels; the bin-local design does not move those bytes (traffic = ncu DRAM bytes "
                "of the same kernels per step), it is bound by instruction issue and shared-memory atomics (DESIGN 4); "
                "gups_* = dependent 16 B load + red.add on random 32 B sectors of an 8 GiB table, measured in this run",
    }
    cpu = None
    if world == 1 and not os.environ.get("MFKC_BENCH_NO_CPU"):
        try:
            import __graft_entry__ as g
            threads = os.cpu_count() or 1
            kmers, dt, kind = cpu_baseline_run(CPU_SAMPLE_READS, threads)
            cpu = {"value": kmers / dt, "unit": "kmers/s", "cores": threads, "kind": kind,
                   "sample": ("first %d of the %d reads as FASTQ through the reference JVM (kmer-counter-many -p %d, parsing included)"
                              % (CPU_SAMPLE_READS, N_READS, threads)) if kind == "reference" else
                             ("first %d of the %d reads (oracle/ref_cpu.c, C restatement of the reference's threaded "
                              "algorithm; the reference JVM cannot run here)" % (CPU_SAMPLE_READS, N_READS))}
        except Exception as ex:                                   # the checker is optional for the measurement itself
            cpu = {"value": None, "unit": "kmers/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}

    ingest = None
    if world == 1 and not os.environ.get("MFKC_BENCH_NO_INGEST"):
        try:
            ingest = host_ingest_numbers()
        except Exception as ex:                                   # informational only
            ingest = {"failed": repr(ex)}

    line = {
        "metric": "canonical 31-mers counted/s", "value": value, "unit": "kmers/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": "kmers/s", "h2d_bytes_per_step": int(n_bases + (n_reads + 1) * 8),
                "d2h_bytes_per_step": int(out_bytes + 32768 * 8), "ms_per_step": None if e2e_failed else 1e3 * e2e_s / args.steps,
                "error": (e2e_error or "the end-to-end pass failed on another rank") if e2e_failed else None,
                "host_to_device_gbs_all_gpus": link_gbs, "h2d_floor_ms_per_step": 1e3 * n_bases * world / (link_gbs * 1e9),
                "mode": "%d contexts take the samples in turn (H2D of sample i+1 overlaps count + emit + D2H of sample i)" % n_lanes if pipelined
                        else "one sample after the other"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "verified": verified,
        "host_ingest": ingest,
        "result": {"kmers_per_step_per_gpu": kmers_per_step, "distinct": st["distinct"], "records_gt_b": int(n_good), "bins": kc.bin_stats(),
                   "host_wall_ms_per_step": 1e3 * wall / args.steps,
                   "host_phase_ms_per_step": {k_: 1e3 * v / args.steps for k_, v in phases.items()}},
    }
    print(json.dumps(line))
    kc.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
