#!/usr/bin/env python
"""Per-rank host timeline of the sharded (peer-memory) step; run under torchrun.
MFKC_BENCH_READS = reads per GPU, MFKC_BENCH_K = k (k = 55: the 128-bit layout of BASELINE config 5)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import metafast_b200 as m
from metafast_b200.sharded import P2PShardedStep, exchange_table

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N, L, K, B = int(os.environ.get("MFKC_BENCH_READS", 20_000_000)), 150, int(os.environ.get("MFKC_BENCH_K", 31)), 1_000_000
kc = m.KmerCounter(K, device=lr, expected_kmers=N * (L - K + 1), n_shards=world, shard_id=rank)
cfg = m.synth_cfg(sample=rank)
d_b = kc.device_alloc(N * L); d_o = kc.device_alloc((N + 1) * 8)
kept = C.c_uint64()
kc._ck(kc.lib.mfkc_synth_reads_device(kc.h, C.byref(cfg), rank * N, N, C.c_void_p(d_b), C.c_void_p(d_o), C.byref(kept)))
n = kept.value
sh = P2PShardedStep(kc, dist, world, rank, B, L, K, N)
for it in range(4):
    kc.sync(); dist.barrier(); torch.cuda.synchronize()
    t = [time.perf_counter()]
    kc.reset(); t.append(time.perf_counter())
    sh.begin(); t.append(time.perf_counter())
    for s in range(0, n, B):
        e = min(n, s + B); kc.p2p_extract(d_b + s * L, d_o + s * 8, e - s, (e - s) * L)
    t.append(time.perf_counter())
    counts = kc.p2p_counts(world); t.append(time.perf_counter())
    rows = exchange_table(dist, [[counts[d]] for d in range(world)], device="cuda"); t.append(time.perf_counter())
    kc.p2p_drain(sum(r[0] for r in rows)); t.append(time.perf_counter())
    kc.flush(); t.append(time.perf_counter())
    ng = kc.emit_begin(2); t.append(time.perf_counter())
    names = ["reset", "begin", "extract(issue)", "counts(sync)", "exchange", "drain(issue)", "flush", "emit_begin"]
    st = kc.stats()
    print("rank %d it %d total %.1f ms | " % (rank, it, 1e3 * (t[-1] - t[0])) + "  ".join("%s %.1f" % (a, 1e3 * (t[i + 1] - t[i])) for i, a in enumerate(names))
          + " | in %.0fM distinct %.0fM good %.0fM" % (sum(r[0] for r in rows) / 1e6, st["distinct"] / 1e6, ng / 1e6), flush=True)
dist.destroy_process_group()
