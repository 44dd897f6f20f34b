"""N>1 host logic on CPU: world_size 2, gloo backend (the GPU kernels are covered by
tests/test_gpu_parity.py::test_logical_shards_equal_unsharded and by bench.py --gpus N)."""
import os
import socket

import numpy as np

from oracle import oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_exchange_and_merge(built, tmp_path):
    import torch.multiprocessing as mp
    from tests import _gloo_worker
    import metafast_b200 as m
    from metafast_b200 import sharded
    world = 2
    mp.spawn(_gloo_worker.run, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [open(os.path.join(tmp_path, "shard%d.bin" % r), "rb").read() for r in range(world)]
    merged = sharded.merge_sorted_records(parts)
    cfg = m.synth_cfg(total_genome_bp=60000, n_genomes=3, n_read_ppm=0, read_len=100)
    reads = [bytes(r).decode() for r in m.synth_reads_host(cfg, 0, 1500)]
    counts = orc.count_reads(reads, 21)
    assert merged == orc.kmers_bin(counts, 1, 21)                # identical to the unsharded result
    assert all(len(p) > 0 for p in parts)
    hist = np.load(os.path.join(tmp_path, "hist.npy"))
    assert {int(c): int(hist[c]) for c in np.nonzero(hist)[0]} == orc.histogram(counts)


def test_peer_memory_protocol_two_ranks(built, tmp_path):
    """P2PShardedStep's host protocol (handle exchange -> attach, staging reset behind a barrier, per-owner totals as
    the barrier before the drain) with 2 gloo ranks and a stub in place of the CUDA context."""
    import torch.multiprocessing as mp
    from tests import _gloo_worker
    import metafast_b200 as m
    from metafast_b200 import sharded
    world = 2
    mp.spawn(_gloo_worker.run_p2p, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [open(os.path.join(tmp_path, "p2p_shard%d.bin" % r), "rb").read() for r in range(world)]
    cfg = m.synth_cfg(total_genome_bp=60000, n_genomes=3, n_read_ppm=0, read_len=100)
    reads = [bytes(r).decode() for r in m.synth_reads_host(cfg, 0, 1200)]
    assert sharded.merge_sorted_records(parts) == orc.kmers_bin(orc.count_reads(reads, 21), 1, 21)
    log2, seg_cap = sharded.p2p_geometry(20_000_000 * 120, 8)
    assert 8 <= log2 <= 14 and (8 << log2) * seg_cap * 16 < 20e9         # cfg2 at 8 GPUs: staging fits comfortably


def test_lanes_of_sharded_steps_do_not_wait_in_a_cycle(built, tmp_path):
    """Several samples in flight per rank (bench.py's end-to-end path for N > 1): the exchange of a lane waits for the
    same lane of the other ranks, so it must run outside the rank-local lock that orders the lanes' submissions.
    With the exchange under the lock this test ends in the groups' timeout."""
    import torch.multiprocessing as mp
    from tests import _gloo_worker
    import metafast_b200 as m
    world = 3
    mp.spawn(_gloo_worker.run_p2p_lanes, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [tuple(int(x) for x in open(os.path.join(tmp_path, "lanes_rank%d.txt" % r)).read().split()) for r in range(world)]
    cfg = m.synth_cfg(total_genome_bp=20000, n_genomes=2, n_read_ppm=0, read_len=100)
    reads = [bytes(r).decode() for r in m.synth_reads_host(cfg, 0, 60 * world)]
    counts = orc.count_reads(reads, 21)
    assert sum(g[0] for g in got) == len(counts) and sum(g[1] for g in got) == sum(counts.values())
