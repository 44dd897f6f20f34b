"""A stand-in for metafast_b200 with the device taken out, for ONE purpose: running bench.py's control flow on a machine
without a GPU (tests/test_bench_dry_run.py).  "Device memory" is host memory, the "kernels" are the C oracle
(oracle/ref_cpu.c through tests/_oracle_c.py), times are made up.  Test infrastructure: nothing under metafast_b200/ or
bench.py imports this, and no number that comes out of it means anything -- only that the script runs through, calls
the API in a legal order and prints a well-formed line."""
import ctypes as C
import hashlib
import threading
import types

import numpy as np

import metafast_b200 as real
from metafast_b200 import _abi
from tests import _oracle_c

_mem = {}                      # address -> numpy array (keeps the "device" allocations alive)
_lock = threading.Lock()
_answers = {}                  # (reads, k, b, shard) -> (records, histogram, distinct)


def _view(addr, nbytes):
    return np.frombuffer((C.c_uint8 * nbytes).from_address(addr), dtype=np.uint8)


class FakeLib:
    """the real library for everything host-side; the two device entry points bench.py calls directly are emulated"""

    def __init__(self):
        self._real = real.load()

    def __getattr__(self, name):
        return getattr(self._real, name)

    def mfkc_device_count(self):
        return 1

    def mfkc_synth_reads_device(self, h, cfg_ref, first, n, d_b, d_o, kept_ref):
        cfg = cfg_ref._obj
        raw = real.synth_reads_host(cfg, int(first), int(n))
        keep = ~(raw == ord("N")).any(axis=1)
        rows = np.ascontiguousarray(raw[keep])
        nk, L = rows.shape
        _view(d_b.value, nk * L)[:] = rows.reshape(-1)
        offs = np.arange(nk + 1, dtype=np.uint64) * np.uint64(L)
        _view(d_o.value, (nk + 1) * 8)[:] = offs.view(np.uint8)
        kept_ref._obj.value = nk
        return 0


class FakeCounter:
    calls = []                 # (instance, method, ...) log over all instances: the test checks the order of the API calls
    serial = 0

    def __init__(self, k, min_seq_len=0, device=0, variant=_abi.VARIANT_HASH, table_slots=0, expected_distinct=0, n_shards=0,
                 shard_id=0, max_table_bytes=0, staging_bytes=0, region_shift=0, expected_kmers=0):
        self.n_shards, self.shard_id = max(int(n_shards), 1), int(shard_id)
        self.k, self.variant, self.rec_size = k, variant, 10
        self.lib, self.h = FakeLib(), C.c_void_p(1)
        self.closed = False
        with _lock:
            FakeCounter.serial += 1
            self.serial = FakeCounter.serial
        self._prof_on, self._prof = False, {}
        self._new_sample()
        self._pinned = []

    # ---- life cycle
    def _new_sample(self):
        self._parts, self._flushed, self._result, self._emitted = [], False, None, None
        self._p2p_sample, self._drained, self._counted = False, None, False

    def _ck(self, rc):
        assert rc == 0, rc

    def close(self):
        self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _log(self, *what):
        assert not self.closed, "call on a closed context"
        with _lock:
            FakeCounter.calls.append((self.serial,) + what)

    def _launch(self, name, n=1):
        if self._prof_on:
            ms, cnt = self._prof.get(name, (0.0, 0))
            self._prof[name] = (ms + 0.001 * n, cnt + n)

    # ---- raw device helpers
    def device_alloc(self, nbytes):
        a = np.zeros(max(int(nbytes), 1), dtype=np.uint8)
        with _lock:
            _mem[a.ctypes.data] = a
        return a.ctypes.data

    def device_free(self, ptr):
        with _lock:
            _mem.pop(ptr)

    def h2d(self, dptr, arr):
        _view(dptr, arr.nbytes)[:] = arr.reshape(-1).view(np.uint8)

    def d2h(self, arr, dptr):
        arr.reshape(-1).view(np.uint8)[:] = _view(dptr, arr.nbytes)

    def sync(self):
        self._log("sync")

    def pinned(self, nbytes, dtype=np.uint8):
        a = np.zeros(int(nbytes), dtype=np.uint8)
        self._pinned.append(a)
        return a.view(dtype)

    def timer_start(self):
        self._t0 = True

    def timer_stop_ms(self):
        assert self._t0
        return 12.5

    def profile(self, enable=None, reset=False):
        if enable is not None:
            self._prof_on = bool(enable)
        if reset:
            self._prof = {}
        return dict(self._prof)

    def gups(self, nbytes, n_updates, mode=1, window_bytes=0, blocks_per_window=0):
        return 17.0

    # ---- the counter
    def reset(self):
        self._log("reset")
        self._new_sample()

    def _add(self, bases, offs):
        assert not self._flushed, "submit after flush without reset"
        offs = np.asarray(offs, dtype=np.uint64)
        b0, b1 = int(offs[0]), int(offs[-1])
        self._parts.append((np.array(bases[b0:b1], copy=True), offs - offs[0]))
        self._launch("mark_read_ends"); self._launch("extract_skm_shard" if self._p2p_sample else "extract_skm")

    def submit(self, bases, offsets):
        self._log("submit", len(offsets) - 1)
        assert bases.dtype == np.uint8 and offsets.dtype == np.uint64
        self._add(bases, offsets)

    def submit_device(self, d_bases, d_offsets, n_reads, n_bases):
        self._log("submit_device", n_reads)
        offs = _view(d_offsets, (n_reads + 1) * 8).view(np.uint64).copy()
        assert int(offs[-1] - offs[0]) == n_bases
        self._add(_view(d_bases, n_bases), offs - offs[0])       # like the C side: offsets relative to offsets[0], pointer already advanced

    def flush(self):
        self._log("flush")
        self._flushed = True

    @staticmethod
    def _join(parts):
        if not parts:
            return np.zeros(1, dtype=np.uint8), np.zeros(1, dtype=np.uint64)
        bases = np.concatenate([p[0] for p in parts])
        offs, run = [np.zeros(1, dtype=np.uint64)], 0
        for pb, po in parts:
            offs.append(po[1:] + np.uint64(run))
            run += len(pb)
        return bases, np.concatenate(offs)

    def _count(self, b):
        if self._result is None or self._result[0] != b:
            bases, offsets = self._join(self._drained if self._p2p_sample else self._parts)
            n_kmers = int(sum(int(offsets[i + 1] - offsets[i]) - self.k + 1 for i in range(len(offsets) - 1)
                              if int(offsets[i + 1] - offsets[i]) >= self.k))
            # the steps of a bench run count the same reads over and over: remember the answers
            key = (hashlib.sha1(bases.tobytes()).hexdigest(), hashlib.sha1(offsets.tobytes()).hexdigest(), self.k, b, self.n_shards, self.shard_id)
            with _lock:
                hit = _answers.get(key)
            if hit is None:
                if self.n_shards == 1:
                    rec, hist, distinct, _ = _oracle_c.count(bases, offsets, self.k, b, P=2)
                else:                                              # this shard's hash range of what all ranks staged
                    rec0, _, _, _ = _oracle_c.count(bases, offsets, self.k, 0, P=2)
                    r = np.frombuffer(rec0, dtype=np.uint8).reshape(-1, 10)
                    keys = r[:, :8].copy().view(">u8").reshape(-1).astype(np.uint64)
                    cnts = r[:, 8:].copy().view(">i2").reshape(-1).astype(np.int64)
                    mine = np.array([self.lib.mfkc_owner_shard(int(x), self.n_shards) == self.shard_id for x in keys], dtype=bool) \
                        if len(keys) else np.zeros(0, dtype=bool)
                    hist = np.bincount(cnts[mine], minlength=32768).astype(np.uint64)
                    distinct = int(mine.sum())
                    rec = r[mine & (cnts > b)].tobytes()
                hit = (rec, hist, distinct)
                with _lock:
                    _answers[key] = hit
            self._result = (b, hit[0], hit[1], hit[2], n_kmers)
        return self._result

    # ---- the peer-memory exchange: "staging buffers" are files every rank can read (MFKC_FAKE_SHARE_DIR)
    def p2p_stage_create(self, log2_buckets, seg_cap):
        raise AssertionError("the dry run covers the bin-local exchange")

    def p2p_stage_create_bins(self, bins_per_shard, seg_cap, ovf_cap):
        assert bins_per_shard >= 16 and seg_cap > 0 and ovf_cap > 0
        self._stage_id = "%d_%d" % (self.shard_id, self.serial)
        self._peers, self._sample_no = {}, 0

    def p2p_export(self):
        return self._stage_id.encode().ljust(128, b".")

    def p2p_attach(self, rank, handles):
        assert (handles is None) == (rank == self.shard_id)
        self._peers[rank] = self._stage_id if handles is None else handles.rstrip(b".").decode()

    def p2p_stage_reset(self):
        self._log("p2p_stage_reset")
        assert sorted(self._peers) == list(range(self.n_shards)), "staging used before every peer is attached"
        self._sample_no += 1
        self._p2p_sample, self._parts, self._drained, self._counted = True, [], None, False

    def _stage_file(self, stage_id):
        import os
        return os.path.join(os.environ["MFKC_FAKE_SHARE_DIR"], "stage_%s_%d.npz" % (stage_id, self._sample_no))

    def p2p_extract(self, d_bases, d_offsets, n_reads, n_bases):
        self._log("p2p_extract", n_reads)
        assert self._p2p_sample and not self._counted
        offs = _view(d_offsets, (n_reads + 1) * 8).view(np.uint64).copy()
        self._add(_view(d_bases, n_bases), offs - offs[0])

    def p2p_submit(self, bases, offsets):
        self._log("p2p_submit", len(offsets) - 1)
        assert self._p2p_sample and not self._counted
        self._add(bases, offsets)

    def p2p_counts(self, n_shards):
        self._log("p2p_counts")
        bases, offsets = self._join(self._parts)
        np.savez(self._stage_file(self._stage_id), bases=bases, offsets=offsets)     # "all my records are in my staging buffer"
        self._counted = True
        n = int(sum(max(0, int(offsets[i + 1] - offsets[i]) - self.k + 1) for i in range(len(offsets) - 1)))
        return [n // n_shards + (1 if d < n % n_shards else 0) for d in range(n_shards)]     # made-up split, exact total

    def p2p_drain(self, n_kmers_in):
        self._log("p2p_drain")
        assert self._counted, "drain before the totals were exchanged"
        self._drained = []
        for r in range(self.n_shards):                             # read "peer memory"
            z = np.load(self._stage_file(self._peers[r]))
            self._drained.append((z["bases"], z["offsets"]))
        self._launch("bin_count")

    def emit_begin(self, threshold):
        self._log("emit_begin")
        assert self._flushed, "emit_begin before flush"
        r = self._count(threshold)
        self._emitted = r[1]
        self._launch("bin_count"); self._launch("radix_sort", 5); self._launch("records")
        return len(r[1]) // 10

    def emit(self, threshold, chunk_bytes=16777200):
        self.emit_begin(threshold)
        return self._emitted

    def emit_into(self, threshold, out):
        n = self.emit_begin(threshold) * 10
        if n > out.nbytes:
            raise ValueError("output buffer too small")
        out[:n] = np.frombuffer(self._emitted, dtype=np.uint8)
        return n

    def histogram(self):
        self._log("histogram")
        assert self._flushed
        return self._count(self._result[0] if self._result else 0)[2].copy()

    def stats(self):
        assert self._flushed
        r = self._count(self._result[0] if self._result else 0)
        return {"distinct": int(r[3]), "kmers": r[4], "total_seq": 0, "good_seq": 0, "total_len": 0, "good_len": 0}

    def bin_stats(self):
        return dict(zip(("bin_mode", "bins", "seg_cap", "heavy_entries", "heavy_recs", "split_passes", "overflow_recs", "staged_recs"), [1] + [0] * 7))


def module():
    """a module object that looks like metafast_b200 to bench.py"""
    m = types.ModuleType("metafast_b200")
    for name in dir(real):
        if not name.startswith("__"):
            setattr(m, name, getattr(real, name))
    m.__path__ = real.__path__                 # so that `from metafast_b200.sharded import ...` still resolves
    m.KmerCounter = FakeCounter
    lib = FakeLib()
    m.load = lambda: lib
    return m
