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
