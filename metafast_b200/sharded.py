"""Hash-range sharding of the k-mer space over the GPUs of one box (one process per GPU).

No reference analogue (the reference is one JVM).  Layout:

    reads        -> split between ranks (any split is valid: counting is order-free)
    keys         -> owner(key) = mfkc_owner_shard(key, G): a hash independent of the table hash
    per batch    -> rank r: extract super-k-mer records bucketed by owner (mfkc_skm_extract_bucketed, CUDA)
                    all-to-all of the bucket sizes, then of the records (NCCL over NVLink)
                    owner: file the records under its table regions (mfkc_skm_count_device, CUDA)
                    (a key-at-a-time flavour exists too: mfkc_extract_bucketed / mfkc_count_keys_device)
    results      -> per-shard histogram summed; per-shard key-sorted records merged by key
                    (shards hold disjoint key sets, so the merge is an interleave)

torch.distributed is only plumbing here (rendezvous + NCCL); every device kernel is libmfkc's.
The exchange helpers are backend-agnostic so that the host logic is testable with gloo on CPU.
"""
from __future__ import annotations

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


def merge_sorted_records(parts: Sequence[bytes], record_size: int = 10, threads: int = 0) -> bytes:
    """Deterministic k-way merge by key of per-shard key-sorted record streams: libmfkc's native, multi-threaded
    mfkc_merge_records (host C++; the shards own disjoint key sets, so the merge is an interleave)."""
    import ctypes as C
    from . import _abi
    lib = _abi.load()
    arrs = [np.frombuffer(p, dtype=np.uint8) if len(p) else np.zeros(1, dtype=np.uint8) for p in parts]
    n = np.array([len(p) // record_size for p in parts], dtype=np.uint64)
    ptrs = (C.c_void_p * max(len(parts), 1))(*[a.ctypes.data for a in arrs])
    out = np.empty(int(n.sum()) * record_size or 1, dtype=np.uint8)
    rc = lib.mfkc_merge_records(C.cast(ptrs, C.c_void_p), n.ctypes.data_as(_abi.u64p), len(parts), record_size,
                                out.ctypes.data_as(C.c_void_p), threads)
    if rc != 0:
        raise _abi.MfkcError(rc, "mfkc_merge_records")
    return out[: int(n.sum()) * record_size].tobytes()


def exchange_table(dist, rows: Sequence[Sequence[int]], device=None, group=None):
    """all-to-all of one small int64 row per destination -> one row per source.  `group`: a process group of its own
    (CPU / gloo) when several sample lanes of one rank exchange concurrently from different threads."""
    import torch
    world = dist.get_world_size()
    width = len(rows[0])
    if group is None and dist.get_backend() == "nccl":
        send = torch.tensor([list(r) for r in rows], dtype=torch.int64, device=device).reshape(world, width)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        return [[int(x) for x in row] for row in recv.tolist()]
    send = torch.tensor([list(r) for r in rows], dtype=torch.int64).reshape(world, width)
    gathered = [torch.empty(world, width, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, send, group=group)
    me = dist.get_rank()
    return [[int(x) for x in gathered[src][me].tolist()] for src in range(world)]


class ShardedStep:
    """Per-rank driver of the sharded counting pass used by bench.py (N > 1).

    Every round: extract super-k-mer records bucketed by owner shard (CUDA) -> exchange the
    per-destination sizes -> all-to-all of the records (NCCL over NVLink, 16 B per ~6 k-mers) ->
    file the received records under their table regions (CUDA).  The region-blocked drain and
    the emit run per shard afterwards, exactly as on one GPU."""

    def __init__(self, kc, dist, world: int, rank: int, batch_reads: int, read_len: int, k: int):
        import torch
        self.kc, self.dist, self.world, self.rank = kc, dist, world, rank
        self.batch_reads, self.read_len, self.k = batch_reads, read_len, k
        self.torch = torch
        kmers = batch_reads * (read_len - k + 1)
        # ~5-6 k-mers per record on Illumina-like reads; 3x slack per destination, retry in halves beyond
        self.seg_cap = int(3.0 * kmers / 5 / world) + 65536
        self.send = torch.empty(world * self.seg_cap * 2, dtype=torch.int64, device="cuda")     # 2 x int64 = one record
        # two receive buffers: the restage kernel of round r (side stream) reads one while the all-to-all of
        # round r+1 fills the other and the extraction of round r+1 runs on the main stream
        self.recv2 = [torch.empty(world * self.seg_cap * 2, dtype=torch.int64, device="cuda") for _ in range(2)]
        self.round = 0
        self.stage_b = torch.empty(batch_reads * read_len + 64, dtype=torch.uint8, device="cuda")
        self.stage_o = torch.empty(batch_reads + 1, dtype=torch.int64, device="cuda")

    def _exchange(self, rec_counts, kmer_counts, overflow):
        """one exchange round; returns True when some rank overflowed a segment (nothing was used)"""
        kc, torch, dist, world = self.kc, self.torch, self.dist, self.world
        rows = exchange_table(dist, [[rec_counts[d], kmer_counts[d], 1 if overflow else 0] for d in range(world)], device="cuda")
        if any(r[2] for r in rows):
            return True
        recv = self.recv2[self.round & 1]
        self.round += 1
        kc.skm_count_wait()                        # the restage kernels that read the receive buffers are done
        ins, outs, off = [], [], 0
        for d in range(world):
            ins.append(self.send[d * self.seg_cap * 2: d * self.seg_cap * 2 + 2 * rec_counts[d]])
        for src in range(world):
            outs.append(recv[off: off + 2 * rows[src][0]])
            off += 2 * rows[src][0]
        dist.all_to_all(outs, ins)
        torch.cuda.current_stream().synchronize()  # records have landed before libmfkc's stream reads them
        n_recs = sum(r[0] for r in rows)
        if n_recs:
            kc.skm_count_device(recv.data_ptr(), n_recs, sum(r[1] for r in rows))
        return False

    def _batch(self, d_bases: int, d_offs: int, n: int, h_offs=None, first=0):
        """reads [first, first+n) of the staged batch; splits in halves when a send segment overflows"""
        kc = self.kc
        if n:
            over, rc, kmc = kc.skm_extract_bucketed(d_bases + first * self.read_len, d_offs + first * 8, n, n * self.read_len,
                                                    self.send.data_ptr(), self.seg_cap, self.world)
        else:
            over, rc, kmc = False, [0] * self.world, [0] * self.world
        if self._exchange(rc, kmc, over):
            half = n // 2                          # every rank splits (also those that did not overflow): rounds stay aligned
            self._batch(d_bases, d_offs, half, h_offs, first)
            self._batch(d_bases, d_offs, n - half, h_offs, first + half)

    def _rounds(self, n_reads: int) -> int:
        """Ranks may hold slightly different numbers of reads (N-reads are dropped per rank): agree on
        the maximum number of exchange rounds up front; ranks that run out send empty buckets."""
        torch, dist = self.torch, self.dist
        t = torch.tensor([(n_reads + self.batch_reads - 1) // self.batch_reads], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())

    def run_device(self, d_bases: int, d_offs: int, n_reads: int):
        for r in range(self._rounds(n_reads)):
            s = r * self.batch_reads
            e = min(n_reads, s + self.batch_reads)
            self._batch(d_bases + s * self.read_len, d_offs + s * 8, max(0, e - s))

    def run_host(self, h_bases: np.ndarray, h_offs: np.ndarray, n_reads: int):
        for r in range(self._rounds(n_reads)):
            s = r * self.batch_reads
            e = min(n_reads, s + self.batch_reads)
            if e > s:
                b0, b1 = int(h_offs[s]), int(h_offs[e])
                self.kc.h2d(self.stage_b.data_ptr(), h_bases[b0:b1])
                self.kc.h2d(self.stage_o.data_ptr(), h_offs[s:e + 1])
            self._batch(self.stage_b.data_ptr(), self.stage_o.data_ptr(), max(0, e - s))


def p2p_geometry(kmers_per_rank: int, world: int, table_bytes_per_rank: int = 0):
    """(log2 coarse buckets, records per segment) of the peer-memory staging: a bucket's slice of the owner's
    table (table / buckets) should be a few MiB so that it stays L2-resident while it is drained; segments get
    2x the expected fill (~5 k-mers per record on Illumina-like reads; the minimizer skew averages out over the
    ~500 minimizers of a segment)."""
    tb = table_bytes_per_rank or 16 * 3 * max(kmers_per_rank // 4, 1)
    log2 = 4
    while log2 < 14 and (tb >> log2) > (8 << 20):
        log2 += 1
    seg_cap = int(2.0 * kmers_per_rank / 5 / (world << log2)) + 256
    return log2, seg_cap


def p2p_bin_geometry(kmers_per_rank: int, world: int, k: int, instances_per_distinct: float = 0.0, slack: float = 0.0):
    """(bins per shard, records per segment, overflow-list records) of the bin-local peer-memory staging:
    libmfkc's mfkc_p2p_bin_geometry (host arithmetic; every rank computes the same numbers)."""
    import ctypes as C
    from . import _abi
    bins, seg, ovf = C.c_uint32(), C.c_uint64(), C.c_uint64()
    rc = _abi.load().mfkc_p2p_bin_geometry(kmers_per_rank, world, k, instances_per_distinct, slack, C.byref(bins), C.byref(seg), C.byref(ovf))
    if rc != 0:
        raise _abi.MfkcError(rc, "mfkc_p2p_bin_geometry")
    return bins.value, seg.value, ovf.value


class P2PShardedStep:
    """Per-rank driver of the sharded counting pass over peer memory (default of bench.py for N > 1).

    Every rank stages its super-k-mer records in its own HBM, bucketed by (owner, coarse bucket); after the
    (tiny) exchange of the per-owner k-mer totals -- which doubles as the barrier -- every owner drains its
    segments straight out of all peers' staging buffers over NVLink inside the counting kernel
    (mfkc_p2p_drain -> drain_p2p_kernel).  torch.distributed only moves the 128-byte IPC handles and the
    per-owner totals."""

    def __init__(self, kc, dist, world: int, rank: int, batch_reads: int, read_len: int, k: int, reads_per_rank: int, group=None):
        import torch
        self.kc, self.dist, self.world, self.rank, self.torch = kc, dist, world, rank, torch
        self.group = group                 # a CPU process group of this lane's own (several lanes per rank run from threads)
        self.batch_reads, self.read_len, self.k = batch_reads, read_len, k
        kmers = reads_per_rank * (read_len - k + 1)
        from . import _abi
        self.bins = k <= 31 and getattr(kc, "variant", _abi.VARIANT_HASH) == _abi.VARIANT_HASH
        if self.bins:                  # bin-local count straight out of the peers' staging buffers
            self.n_bins, self.seg_cap, self.ovf_cap = p2p_bin_geometry(kmers, world, k)
            kc.p2p_stage_create_bins(self.n_bins, self.seg_cap, self.ovf_cap)
        else:                          # region-blocked table drained from peer memory (k > 31, MFKC_VARIANT_HASH_TABLE)
            self.log2, self.seg_cap = p2p_geometry(kmers, world)
            kc.p2p_stage_create(self.log2, self.seg_cap)
        mine = torch.frombuffer(bytearray(kc.p2p_export()), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            mine = mine.cuda()
        handles = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(handles, mine)
        for r in range(world):
            kc.p2p_attach(r, None if r == rank else bytes(handles[r].cpu().numpy().tobytes()))

    def begin(self):
        """call after kc.reset(): every rank has finished draining the previous sample before any staging buffer
        is cleared"""
        self.dist.barrier(group=self.group) if self.group is not None else self.dist.barrier()
        self.kc.p2p_stage_reset()

    def _finish(self):
        kc, world = self.kc, self.world
        counts = kc.p2p_counts(world)                                        # synchronises this rank's extraction
        rows = exchange_table(self.dist, [[counts[d]] for d in range(world)], device="cuda" if self.dist.get_backend() == "nccl" else None,
                              group=self.group)
        kc.p2p_drain(sum(r[0] for r in rows))

    def run_device(self, d_bases: int, d_offs: int, n_reads: int):
        for s in range(0, n_reads, self.batch_reads):
            e = min(n_reads, s + self.batch_reads)
            self.kc.p2p_extract(d_bases + s * self.read_len, d_offs + s * 8, e - s, (e - s) * self.read_len)
        self._finish()

    def submit_host(self, h_bases: np.ndarray, h_offs: np.ndarray, n_reads: int):
        """this rank's part only (host -> device copies + extraction launches): no other rank is waited for, so a caller
        that runs several lanes may hold a rank-local lock around it"""
        for s in range(0, n_reads, self.batch_reads):
            e = min(n_reads, s + self.batch_reads)
            self.kc.p2p_submit(h_bases, h_offs[s:e + 1])

    def finish(self):
        """the exchange of the per-owner totals (every rank of the group takes part) + this owner's count.  Never call it
        while holding a rank-local lock that another lane needs for its own submit_host: two ranks whose lanes took their
        locks in different orders would wait for each other for ever."""
        self._finish()

    def run_host(self, h_bases: np.ndarray, h_offs: np.ndarray, n_reads: int):
        self.submit_host(h_bases, h_offs, n_reads)
        self._finish()


def run_lanes(n_steps: int, lanes, step):
    """Samples 0..n_steps-1 over len(lanes) lanes, one host thread per lane: lane j takes samples j, j+L, j+2L, ...
    step(lane, i) does one sample.  An exception in any lane is re-raised here (after all lanes ended) instead of
    dying with its thread -- a rank must not carry on while its peers wait for the failed lane."""
    import threading
    n_lanes = len(lanes)
    res, errors = [None] * n_steps, []

    def lane(j):
        try:
            for i in range(j, n_steps, n_lanes):
                res[i] = step(lanes[j], i)
        except BaseException as e:          # noqa: BLE001 -- reported below
            errors.append((j, e))
    ts = [threading.Thread(target=lane, args=(j,)) for j in range(n_lanes)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise RuntimeError("lane %d failed: %r" % (errors[0][0], errors[0][1])) from errors[0][1]
    return res
