"""Hash-range sharding of the k-mer space over the GPUs of one box (one process per GPU).

No reference analogue (the reference is one JVM).  Layout:

    reads        -> split between ranks (any split is valid: counting is order-free)
    keys         -> owner(key) = mfkc_owner_shard(key, G): a hash independent of the table hash
    per batch    -> rank r: extract + bucket by owner (mfkc_extract_bucketed, CUDA)
                    all-to-all of the bucket sizes, then of the keys (NCCL over NVLink)
                    owner: count the received keys (mfkc_count_keys_device, CUDA)
    results      -> per-shard histogram summed; per-shard key-sorted records merged by key
                    (shards hold disjoint key sets, so the merge is an interleave)

torch.distributed is only plumbing here (rendezvous + NCCL); every device kernel is libmfkc's.
The exchange helpers are backend-agnostic so that the host logic is testable with gloo on CPU.
"""
from __future__ import annotations

import heapq
from typing import List, Sequence

import numpy as np


def exchange_counts(dist, counts: Sequence[int], device=None):
    """all-to-all of the per-destination bucket sizes -> per-source receive sizes."""
    import torch
    world = dist.get_world_size()
    send = torch.tensor(list(counts), dtype=torch.int64, device=device)
    if dist.get_backend() == "nccl":
        recv = torch.empty(world, dtype=torch.int64, device=device)
        dist.all_to_all_single(recv, send)
        return [int(x) for x in recv.tolist()]
    gathered = [torch.empty(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, send.cpu())
    me = dist.get_rank()
    return [int(gathered[src][me]) for src in range(world)]


def exchange_keys(dist, send, send_counts: Sequence[int], recv, recv_counts: Sequence[int]):
    """Move bucket d of `send` (int64 tensor grouped by destination) to rank d; returns the number
    of keys received into `recv` (grouped by source).  NCCL: one all_to_all_single; gloo (CPU
    tests): point-to-point sends, since gloo has no all-to-all."""
    n_recv = int(sum(recv_counts))
    n_send = int(sum(send_counts))
    if n_recv > recv.numel():
        raise RuntimeError("receive buffer too small: %d > %d" % (n_recv, recv.numel()))
    if dist.get_backend() == "nccl":
        dist.all_to_all_single(recv[:n_recv], send[:n_send], output_split_sizes=list(recv_counts),
                               input_split_sizes=list(send_counts))
        return n_recv
    me, world = dist.get_rank(), dist.get_world_size()
    so = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64)
    ro = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
    reqs = []
    for peer in range(world):
        if peer == me:
            recv[ro[me]:ro[me + 1]] = send[so[me]:so[me + 1]]
            continue
        if send_counts[peer]:
            reqs.append(dist.isend(send[so[peer]:so[peer + 1]].contiguous(), peer))
        if recv_counts[peer]:
            reqs.append(dist.irecv(recv[ro[peer]:ro[peer + 1]], peer))
    for r in reqs:
        r.wait()
    return n_recv


def merge_sorted_records(parts: Sequence[bytes], record_size: int = 10) -> bytes:
    """Deterministic k-way merge by key of per-shard key-sorted record streams (big-endian keys
    compare like the byte strings themselves)."""
    def recs(b):
        return (b[i:i + record_size] for i in range(0, len(b), record_size))
    return b"".join(heapq.merge(*[recs(p) for p in parts]))


class ShardedStep:
    """Per-rank driver of the sharded counting pass used by bench.py (N > 1)."""

    def __init__(self, kc, dist, world: int, rank: int, batch_reads: int, read_len: int, k: int):
        import torch
        self.kc, self.dist, self.world, self.rank = kc, dist, world, rank
        self.batch_reads, self.read_len, self.k = batch_reads, read_len, k
        self.torch = torch
        cap = batch_reads * (read_len - k + 1)
        self.send = torch.empty(cap, dtype=torch.int64, device="cuda")
        self.recv = torch.empty(int(cap * 1.5) + 4096, dtype=torch.int64, device="cuda")
        self.stage_b = torch.empty(batch_reads * read_len + 64, dtype=torch.uint8, device="cuda")
        self.stage_o = torch.empty(batch_reads + 1, dtype=torch.int64, device="cuda")

    def _batch(self, d_bases: int, d_offs: int, n: int):
        kc, torch, dist = self.kc, self.torch, self.dist
        counts = kc.extract_bucketed(d_bases, d_offs, n, n * self.read_len, self.send.data_ptr(), self.send.numel(), self.world)
        rcounts = exchange_counts(dist, counts, device="cuda")
        kc.sync()                                  # the previous batch's count kernel has released self.recv
        n_recv = exchange_keys(dist, self.send, counts, self.recv, rcounts)
        torch.cuda.current_stream().synchronize()  # keys have landed before libmfkc's stream reads them
        kc.count_keys_device(self.recv.data_ptr(), n_recv)

    def _rounds(self, n_reads: int) -> int:
        """Ranks may hold slightly different numbers of reads (N-reads are dropped per rank): agree on
        the maximum number of exchange rounds up front; ranks that run out send empty buckets."""
        torch, dist = self.torch, self.dist
        t = torch.tensor([(n_reads + self.batch_reads - 1) // self.batch_reads], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())

    def _empty_batch(self):
        torch, dist = self.torch, self.dist
        zeros = [0] * self.world
        rcounts = exchange_counts(dist, zeros, device="cuda")
        self.kc.sync()
        n_recv = exchange_keys(dist, self.send, zeros, self.recv, rcounts)
        torch.cuda.current_stream().synchronize()
        if n_recv:
            self.kc.count_keys_device(self.recv.data_ptr(), n_recv)

    def run_device(self, d_bases: int, d_offs: int, n_reads: int):
        for r in range(self._rounds(n_reads)):
            s = r * self.batch_reads
            if s >= n_reads:
                self._empty_batch()
                continue
            e = min(n_reads, s + self.batch_reads)
            self._batch(d_bases + s * self.read_len, d_offs + s * 8, e - s)

    def run_host(self, h_bases: np.ndarray, h_offs: np.ndarray, n_reads: int):
        for r in range(self._rounds(n_reads)):
            s = r * self.batch_reads
            if s >= n_reads:
                self._empty_batch()
                continue
            e = min(n_reads, s + self.batch_reads)
            b0, b1 = int(h_offs[s]), int(h_offs[e])
            self.kc.h2d(self.stage_b.data_ptr(), h_bases[b0:b1])
            self.kc.h2d(self.stage_o.data_ptr(), h_offs[s:e + 1])
            self._batch(self.stage_b.data_ptr(), self.stage_o.data_ptr(), e - s)
