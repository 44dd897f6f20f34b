"""Worker of the world_size-2 gloo test: the N>1 host logic (bucket by owner -> exchange ->
per-shard count -> deterministic merge) with the device kernels replaced by the oracle's
CPU arithmetic.  Exercises metafast_b200.sharded on CPU tensors."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run(rank, world, port, out_dir):
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import metafast_b200 as m
    from metafast_b200 import sharded
    from oracle import oracle as orc
    lib = m.load()
    k, b = 21, 1
    cfg = m.synth_cfg(total_genome_bp=60000, n_genomes=3, n_read_ppm=0, read_len=100)
    n = 1500
    raw = m.synth_reads_host(cfg, 0, n)
    reads = [bytes(r).decode() for r in raw]
    mine = reads[rank::world]                                    # any split of the reads is valid
    counts = {}
    hist = np.zeros(32768, dtype=np.int64)
    rounds = (max(len(reads[r::world]) for r in range(world)) + 399) // 400
    for rd in range(rounds):                                     # several exchange rounds, like the batches of a sample
        batch = mine[rd * 400:(rd + 1) * 400]
        keys = orc.canonical_kmers_np(batch, k) if batch else np.zeros(0, dtype=np.uint64)
        owners = np.array([lib.mfkc_owner_shard(int(x), world) for x in keys], dtype=np.int64)
        order = np.argsort(owners, kind="stable")
        send = torch.from_numpy(keys[order].view(np.int64).copy())
        scounts = [int((owners == d).sum()) for d in range(world)]
        rcounts = sharded.exchange_counts(dist, scounts)
        recv = torch.empty(sum(rcounts) + 8, dtype=torch.int64)
        got = sharded.exchange_keys(dist, send, scounts, recv, rcounts)
        for x in recv[:got].numpy().view(np.uint64):
            x = int(x)
            assert lib.mfkc_owner_shard(x, world) == rank
            counts[x] = min(counts.get(x, 0) + 1, 32767)
    rec = orc.kmers_bin(counts, b, k)
    for c in counts.values():
        hist[c] += 1
    t = torch.from_numpy(hist)
    dist.all_reduce(t)                                           # histogram: sum over shards
    with open(os.path.join(out_dir, "shard%d.bin" % rank), "wb") as f:
        f.write(rec)
    if rank == 0:
        np.save(os.path.join(out_dir, "hist.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


class _StubCounter:
    """Stands in for KmerCounter in the peer-memory protocol test: "staging buffers" are files in a shared directory
    (every rank can read every rank's, like peer HBM), "kernels" are the oracle's CPU arithmetic."""

    def __init__(self, rank, world, k, share_dir, lib):
        self.rank, self.world, self.k, self.dir, self.lib = rank, world, k, share_dir, lib
        self.attached, self.counts, self.log = {}, {}, []

    def p2p_stage_create(self, log2_buckets, seg_cap):
        self.log.append(("create", log2_buckets, seg_cap))

    def p2p_stage_create_bins(self, bins_per_shard, seg_cap, ovf_cap):
        assert bins_per_shard >= 16 and seg_cap > 0 and ovf_cap > 0
        self.log.append(("create", bins_per_shard, seg_cap, ovf_cap))

    def p2p_export(self):
        return (b"rank%03d" % self.rank).ljust(128, b".")

    def p2p_attach(self, r, handle):
        assert (handle is None) == (r == self.rank)
        if handle is not None:
            assert handle == (b"rank%03d" % r).ljust(128, b".")
        self.attached[r] = True

    def p2p_stage_reset(self):
        import numpy as np
        self.staged = [np.zeros(0, dtype=np.uint64) for _ in range(self.world)]

    def p2p_submit(self, bases, offsets):
        import numpy as np
        from oracle import oracle as orc
        reads = [bytes(bases[int(offsets[i]):int(offsets[i + 1])]).decode() for i in range(len(offsets) - 1)]
        keys = orc.canonical_kmers_np(reads, self.k)
        owners = np.array([self.lib.mfkc_owner_shard(int(x), self.world) for x in keys], dtype=np.int64)
        for d in range(self.world):
            self.staged[d] = np.concatenate([self.staged[d], keys[owners == d]])

    def p2p_counts(self, world):
        import numpy as np
        for d in range(world):                                # "every record of this rank is in its staging buffer"
            np.save("%s/stage_%d_to_%d.npy" % (self.dir, self.rank, d), self.staged[d])
        return [len(x) for x in self.staged]

    def p2p_drain(self, n_in):
        import numpy as np
        got = 0
        for src in range(self.world):                         # read "peer memory"
            for x in np.load("%s/stage_%d_to_%d.npy" % (self.dir, src, self.rank)):
                self.counts[int(x)] = min(self.counts.get(int(x), 0) + 1, 32767)
                got += 1
        assert got == n_in                                    # the exchanged totals are exact


def run_p2p(rank, world, port, out_dir):
    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import metafast_b200 as m
    from metafast_b200 import sharded
    from oracle import oracle as orc
    k, b, L = 21, 1, 100
    cfg = m.synth_cfg(total_genome_bp=60000, n_genomes=3, n_read_ppm=0, read_len=L)
    n = 1200
    raw = m.synth_reads_host(cfg, 0, n)
    mine = np.ascontiguousarray(raw[rank::world])
    kc = _StubCounter(rank, world, k, out_dir, m.load())
    step = sharded.P2PShardedStep(kc, dist, world, rank, 250, L, k, len(mine))
    assert sorted(kc.attached) == list(range(world)) and kc.log[0][0] == "create"
    for sample in range(2):                                   # two samples: begin() is the barrier before the staging is cleared
        kc.counts = {}
        step.begin()
        step.run_host(mine.reshape(-1), np.arange(len(mine) + 1, dtype=np.uint64) * np.uint64(L), len(mine))
    with open(os.path.join(out_dir, "p2p_shard%d.bin" % rank), "wb") as f:
        f.write(orc.kmers_bin(kc.counts, b, k))
    dist.barrier()
    dist.destroy_process_group()


def run_p2p_lanes(rank, world, port, out_dir):
    """bench.py's multi-lane end-to-end protocol for N > 1: several samples in flight per rank, one lane = one context +
    one sharded step + one CPU process group; a rank-local lock serialises the lanes' submissions.  Random delays make
    the lanes of different ranks take that lock in different orders."""
    import datetime
    import random
    import threading
    import time
    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import metafast_b200 as m
    from metafast_b200 import sharded
    k, L, n_lanes, n_steps = 21, 100, 3, 24
    cfg = m.synth_cfg(total_genome_bp=20000, n_genomes=2, n_read_ppm=0, read_len=L)
    raw = m.synth_reads_host(cfg, 0, 60 * world)
    mine = np.ascontiguousarray(raw[rank::world])
    offs = np.arange(len(mine) + 1, dtype=np.uint64) * np.uint64(L)
    lib = m.load()
    lanes = []
    for j in range(n_lanes):
        d = os.path.join(out_dir, "lane%d" % j)
        os.makedirs(d, exist_ok=True)
        kc = _StubCounter(rank, world, k, d, lib)
        grp = dist.new_group(backend="gloo", timeout=datetime.timedelta(seconds=240))
        lanes.append((kc, sharded.P2PShardedStep(kc, dist, world, rank, 25, L, k, len(mine), group=grp)))
    link = threading.Lock()
    rnd = random.Random(1000 + rank)
    sizes = []

    def step(lane, i):
        kc, shd = lane
        kc.counts = {}
        shd.begin()
        with link:
            time.sleep(rnd.random() * 0.01)
            shd.submit_host(mine.reshape(-1), offs, len(mine))
        shd.finish()
        time.sleep(rnd.random() * 0.01)
        return len(kc.counts), sum(kc.counts.values())
    res = sharded.run_lanes(n_steps, lanes, step)
    assert len(set(res)) == 1 and res[0][0] > 0                # every sample of every lane: the same counts
    with open(os.path.join(out_dir, "lanes_rank%d.txt" % rank), "w") as f:
        f.write("%d %d\n" % res[0])

    def failing(lane, i):
        raise ValueError("boom")
    try:
        sharded.run_lanes(3, lanes, failing)
        raise AssertionError("run_lanes swallowed a lane's exception")
    except RuntimeError as e:
        assert "boom" in str(e)
    dist.barrier()
    dist.destroy_process_group()


def run_bench_fake_device(rank, world, port, out_dir):
    """bench.py's N > 1 control flow on CPU: one process per "GPU", gloo in place of NCCL, tests/_fake_device.py in place
    of the device (tests/test_bench_dry_run.py)."""
    import contextlib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MFKC_FAKE_SHARE_DIR=out_dir, MFKC_BENCH_NO_INGEST="1", MFKC_BENCH_VERIFY_SHARD_READS="1500",
                      MFKC_BENCH_LANE_TIMEOUT_S="240")
    for k in ("MFKC_BENCH_VARIANT", "MFKC_BENCH_NO_VERIFY", "MFKC_EXCHANGE", "MFKC_BENCH_E2E_SERIAL"):
        os.environ.pop(k, None)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.synchronize = lambda *a, **k: None
    real_init, real_tensor = dist.init_process_group, torch.tensor
    dist.init_process_group = lambda backend=None, **kw: real_init("gloo", rank=rank, world_size=world)
    torch.tensor = lambda *a, **kw: real_tensor(*a, **{k: v for k, v in kw.items() if k != "device"})
    import bench
    import metafast_b200.sharded                                 # noqa: F401  (resolved before the package is swapped)
    from tests import _fake_device
    sys.modules["metafast_b200"] = _fake_device.module()
    bench.N_READS, bench.BATCH_READS, bench.B_THRESHOLD = 3000, 1000, 0
    sys.argv = ["bench.py", "--gpus", str(world), "--steps", "5", "--warmup", "3"]
    with open(os.path.join(out_dir, "bench_rank%d.out" % rank), "w") as f, contextlib.redirect_stdout(f):
        bench.main()
    calls = _fake_device.FakeCounter.calls
    with open(os.path.join(out_dir, "calls_rank%d.txt" % rank), "w") as f:
        f.write("\n".join(" ".join(str(x) for x in c) for c in calls))
