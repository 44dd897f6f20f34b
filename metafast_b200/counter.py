"""Thin Python host layer over the C ABI (used by tests, bench.py and the smoke check).

The classes mirror the seams of the reference that libmfkc replaces:

* ``KmerCounter.submit / flush / emit``  <->  IOUtils.loadReads + IOUtils.printKmers
  (src/io/IOUtils.java:772-803, 45-71)
* ``FeaturesCalculator``                <->  FeaturesCalculatorMain's map set-up,
  IOUtils.calculatePresenceForKmers/Reads and buildAndPrintVector
  (src/tools/FeaturesCalculatorMain.java:97-103,169-236; src/io/IOUtils.java:577-597,806-834)
* ``read_file``                         <->  ReadersUtils.readDnaLazy ([itmo]/io/ReadersUtils.java:85-102)

Everything is executed by libmfkc (CUDA kernels for the device work, C++ for the parsers).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def pack_reads(reads: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """List of read strings -> (ASCII bases uint8[], offsets uint64[n+1])."""
    offsets = np.zeros(len(reads) + 1, dtype=np.uint64)
    if reads:
        offsets[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    bases = np.frombuffer("".join(reads).encode("latin-1"), dtype=np.uint8).copy()
    if bases.size == 0:
        bases = np.zeros(1, dtype=np.uint8)
    return bases, offsets


class _Ctx:
    def __init__(self, k: int, min_seq_len: int = 0, device: int = 0, variant: int = _abi.VARIANT_HASH,
                 table_slots: int = 0, expected_distinct: int = 0, n_shards: int = 0, shard_id: int = 0,
                 max_table_bytes: int = 0, staging_bytes: int = 0, region_shift: int = 0, expected_kmers: int = 0):
        self.lib = _abi.load()
        cfg = _abi.MfkcCfg()
        cfg.struct_size = C.sizeof(_abi.MfkcCfg)
        cfg.k, cfg.min_seq_len, cfg.device, cfg.variant = k, min_seq_len, device, variant
        cfg.n_shards, cfg.shard_id = n_shards, shard_id
        cfg.table_slots, cfg.expected_distinct, cfg.max_table_bytes = table_slots, expected_distinct, max_table_bytes
        cfg.staging_bytes, cfg.region_shift, cfg.expected_kmers = staging_bytes, region_shift, expected_kmers
        h = C.c_void_p()
        rc = self.lib.mfkc_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise _abi.MfkcError(rc, self.lib.mfkc_last_error(None).decode())
        self.h = h
        self.k = k
        self.variant = variant
        self.rec_size = 18 if k > 31 else 10      # 16-byte BE key for 128-bit keys (extension)

    def _ck(self, rc: int):
        if rc != 0:
            raise _abi.MfkcError(rc, self.lib.mfkc_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.mfkc_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- raw device helpers
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._ck(self.lib.mfkc_device_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, ptr: int):
        self._ck(self.lib.mfkc_device_free(self.h, C.c_void_p(ptr)))

    def h2d(self, dptr: int, arr: np.ndarray):
        self._ck(self.lib.mfkc_memcpy_h2d(self.h, C.c_void_p(dptr), _ptr(arr), arr.nbytes))

    def d2h(self, arr: np.ndarray, dptr: int):
        self._ck(self.lib.mfkc_memcpy_d2h(self.h, _ptr(arr), C.c_void_p(dptr), arr.nbytes))

    def sync(self):
        self._ck(self.lib.mfkc_device_sync(self.h))

    def pinned(self, nbytes: int, dtype=np.uint8) -> np.ndarray:
        p = C.c_void_p()
        self._ck(self.lib.mfkc_pinned_alloc(self.h, nbytes, C.byref(p)))
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        a = np.frombuffer(buf, dtype=np.uint8).view(dtype)
        a.flags.writeable = True
        return a

    def timer_start(self):
        self._ck(self.lib.mfkc_timer_start(self.h))

    def timer_stop_ms(self) -> float:
        ms = C.c_float()
        self._ck(self.lib.mfkc_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def profile(self, enable: Optional[bool] = None, reset: bool = False) -> dict:
        if enable is not None:
            self._ck(self.lib.mfkc_profile_enable(self.h, 1 if enable else 0))
        if reset:
            self._ck(self.lib.mfkc_profile_reset(self.h))
        out = {}
        i = 0
        while True:
            name = self.lib.mfkc_profile_name(i)
            if name is None:
                break
            ms, n = C.c_double(), C.c_uint64()
            self._ck(self.lib.mfkc_profile_get(self.h, i, C.byref(ms), C.byref(n)))
            out[name.decode()] = (ms.value, n.value)
            i += 1
        return out

    def gups(self, nbytes: int, n_updates: int, mode: int = 1, window_bytes: int = 0, blocks_per_window: int = 0) -> float:
        """Random-sector microbenchmark (see mfkc_gups_ex); returns milliseconds."""
        ms = C.c_float()
        self._ck(self.lib.mfkc_gups_ex(self.h, nbytes, n_updates, mode, window_bytes, blocks_per_window, C.byref(ms)))
        return ms.value


class KmerCounter(_Ctx):
    """One sample's k-mer counter (BigLong2ShortHashMap + ReadsLoadWorker pool analogue)."""

    def submit(self, bases: np.ndarray, offsets: np.ndarray):
        assert bases.dtype == np.uint8 and offsets.dtype == np.uint64
        self._ck(self.lib.mfkc_submit_reads(self.h, _ptr(bases), _ptr(offsets), len(offsets) - 1))

    def submit_reads(self, reads: Sequence[str]):
        b, o = pack_reads(reads)
        self.submit(b, o)

    def submit_device(self, d_bases: int, d_offsets: int, n_reads: int, n_bases: int):
        self._ck(self.lib.mfkc_submit_reads_device(self.h, C.c_void_p(d_bases), C.c_void_p(d_offsets), n_reads, n_bases))

    def flush(self):
        self._ck(self.lib.mfkc_flush(self.h))

    def reset(self):
        self._ck(self.lib.mfkc_reset(self.h))

    def stats(self) -> dict:
        s = (C.c_uint64 * 6)()
        self._ck(self.lib.mfkc_stats(self.h, s))
        return dict(zip(("distinct", "kmers", "total_seq", "good_seq", "total_len", "good_len"), list(s)))

    def bin_stats(self) -> dict:
        """diagnostics of the bin-local mode (mfkc_bin_stats)"""
        s = (C.c_uint64 * 8)()
        self._ck(self.lib.mfkc_bin_stats(self.h, s))
        return dict(zip(("bin_mode", "bins", "seg_cap", "heavy_entries", "heavy_recs", "split_passes", "overflow_recs", "staged_recs"), list(s)))

    def histogram(self) -> np.ndarray:
        h = np.zeros(_abi.HIST_BINS, dtype=np.uint64)
        self._ck(self.lib.mfkc_histogram(self.h, h.ctypes.data_as(_abi.u64p)))
        return h

    def emit_begin(self, threshold: int) -> int:
        n = C.c_uint64()
        self._ck(self.lib.mfkc_emit_begin(self.h, threshold, C.byref(n)))
        return n.value

    def emit(self, threshold: int, chunk_bytes: int = 16777200) -> bytes:
        """All records (big-endian, ascending key) as one bytes object."""
        n = self.emit_begin(threshold)
        rs = self.rec_size
        out = np.empty(n * rs, dtype=np.uint8) if n else np.empty(0, dtype=np.uint8)
        pos = 0
        w = C.c_size_t()
        while True:
            cap = min(chunk_bytes, out.nbytes - pos)
            self._ck(self.lib.mfkc_emit_next(self.h, C.c_void_p(out.ctypes.data + pos) if cap else None, cap, C.byref(w)))
            if w.value == 0:
                break
            pos += w.value
        assert pos == n * rs
        return out.tobytes()

    def emit_into(self, threshold: int, out: np.ndarray) -> int:
        """Records into a caller-provided (ideally pinned) uint8 buffer; returns bytes written."""
        n = self.emit_begin(threshold)
        rs = self.rec_size
        if n * rs > out.nbytes:
            raise ValueError("output buffer too small: need %d bytes" % (n * rs))
        pos = 0
        w = C.c_size_t()
        while pos < n * rs:
            self._ck(self.lib.mfkc_emit_next(self.h, C.c_void_p(out.ctypes.data + pos), out.nbytes - pos, C.byref(w)))
            if w.value == 0:
                break
            pos += w.value
        return pos

    def emit_device(self) -> Tuple[int, int, int]:
        k, c, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self._ck(self.lib.mfkc_emit_device(self.h, C.byref(k), C.byref(c), C.byref(n)))
        return k.value, c.value, n.value

    # ---- sharding
    def extract_bucketed(self, d_bases: int, d_offsets: int, n_reads: int, n_bases: int, d_keys_out: int,
                         cap_keys: int, n_shards: int) -> List[int]:
        cnt = (C.c_uint64 * max(n_shards, 1))()
        self._ck(self.lib.mfkc_extract_bucketed(self.h, C.c_void_p(d_bases), C.c_void_p(d_offsets), n_reads, n_bases,
                                                C.c_void_p(d_keys_out), cap_keys, cnt))
        return list(cnt)

    def count_keys_device(self, d_keys: int, n: int):
        self._ck(self.lib.mfkc_count_keys_device(self.h, C.c_void_p(d_keys), n))

    def skm_extract_bucketed(self, d_bases: int, d_offsets: int, n_reads: int, n_bases: int, d_recs_out: int,
                             seg_cap: int, n_shards: int):
        """-> (overflowed, records per shard, k-mer instances per shard)"""
        rc_ = (C.c_uint64 * max(n_shards, 1))()
        kc_ = (C.c_uint64 * max(n_shards, 1))()
        rc = self.lib.mfkc_skm_extract_bucketed(self.h, C.c_void_p(d_bases), C.c_void_p(d_offsets), n_reads, n_bases,
                                                C.c_void_p(d_recs_out), seg_cap, rc_, kc_)
        if rc not in (0, 1):
            self._ck(rc)
        return rc == 1, list(rc_), list(kc_)

    def skm_count_device(self, d_recs: int, n_recs: int, n_kmers: int):
        self._ck(self.lib.mfkc_skm_count_device(self.h, C.c_void_p(d_recs), n_recs, n_kmers))

    def skm_count_wait(self):
        self._ck(self.lib.mfkc_skm_count_wait(self.h))

    # ---- peer-memory flavour of the exchange (mfkc_p2p_*)
    def p2p_stage_create(self, log2_buckets: int, seg_cap: int):
        self._ck(self.lib.mfkc_p2p_stage_create(self.h, log2_buckets, seg_cap))

    def p2p_stage_create_bins(self, bins_per_shard: int, seg_cap: int, ovf_cap: int):
        self._ck(self.lib.mfkc_p2p_stage_create_bins(self.h, bins_per_shard, seg_cap, ovf_cap))

    def p2p_export(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        self._ck(self.lib.mfkc_p2p_export(self.h, buf))
        return bytes(buf)

    def p2p_attach(self, rank: int, handles: Optional[bytes]):
        buf = (C.c_uint8 * 128).from_buffer_copy(handles) if handles is not None else None
        self._ck(self.lib.mfkc_p2p_attach(self.h, rank, buf))

    def p2p_attach_ctx(self, rank: int, peer: "KmerCounter"):
        self._ck(self.lib.mfkc_p2p_attach_ctx(self.h, rank, peer.h))

    def p2p_stage_reset(self):
        self._ck(self.lib.mfkc_p2p_stage_reset(self.h))

    def p2p_extract(self, d_bases: int, d_offsets: int, n_reads: int, n_bases: int):
        self._ck(self.lib.mfkc_p2p_extract(self.h, C.c_void_p(d_bases), C.c_void_p(d_offsets), n_reads, n_bases))

    def p2p_submit(self, bases: np.ndarray, offsets: np.ndarray):
        self._ck(self.lib.mfkc_p2p_submit_reads(self.h, _ptr(bases), _ptr(offsets), len(offsets) - 1))

    def p2p_counts(self, n_shards: int) -> List[int]:
        out = (C.c_uint64 * max(n_shards, 1))()
        self._ck(self.lib.mfkc_p2p_counts(self.h, out))
        return list(out)

    def p2p_drain(self, n_kmers_in: int):
        self._ck(self.lib.mfkc_p2p_drain(self.h, n_kmers_in))


class KmerSet:
    """One (k-mer -> short) map of the .kmers.bin set-algebra tools on the device (mfkc_kset_*): the
    BigLong2ShortHashMap of src/tools/KmersFilter.java, UniqueKmersMultipleSamplesFinder.java and
    KmersSamplesCounter.java.  `ctx` is any KmerCounter / FeaturesCalculator context (device + streams)."""
    ADD, INC, ZERO = 0, 1, 2

    def __init__(self, ctx: "_Ctx"):
        self.ctx, self.lib = ctx, ctx.lib
        h = C.c_void_p()
        ctx._ck(self.lib.mfkc_kset_create(ctx.h, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.mfkc_kset_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @classmethod
    def load(cls, ctx: "_Ctx", datas: Sequence[bytes], threshold: int, chunk: int = 16777200) -> "KmerSet":
        """IOUtils.loadKmers(files, threshold) (src/io/IOUtils.java:369-401)"""
        ks = cls(ctx)
        for d in datas:
            a = np.frombuffer(d, dtype=np.uint8)
            for s in range(0, len(d), chunk):
                part = a[s:s + chunk]
                ctx._ck(ks.lib.mfkc_kset_load_records(ks.h, _ptr(part), part.nbytes // 10, threshold))
        ctx._ck(ks.lib.mfkc_kset_load_finish(ks.h))
        return ks

    def size(self) -> int:
        n = C.c_uint64()
        self.ctx._ck(self.lib.mfkc_kset_size(self.h, C.byref(n)))
        return n.value

    def reset_values(self):
        self.ctx._ck(self.lib.mfkc_kset_reset_values(self.h))

    def update(self, src: "KmerSet", op: int, thr: int):
        self.ctx._ck(self.lib.mfkc_kset_update(self.h, src.h, op, thr))

    def select(self, filt: Optional["KmerSet"], threshold: int, filter_threshold: int = 0, chunk: int = 16777200) -> bytes:
        """IOUtils.filterAndPrintKmers (src/io/IOUtils.java:101-123); records in ascending key order"""
        n = C.c_uint64()
        self.ctx._ck(self.lib.mfkc_kset_select_begin(self.h, filt.h if filt is not None else None, threshold, filter_threshold, C.byref(n)))
        out = np.empty(n.value * 10, dtype=np.uint8)
        pos, w = 0, C.c_size_t()
        while pos < out.nbytes:
            self.ctx._ck(self.lib.mfkc_kset_select_next(self.h, C.c_void_p(out.ctypes.data + pos), min(chunk, out.nbytes - pos), C.byref(w)))
            if w.value == 0:
                break
            pos += w.value
        assert pos == out.nbytes
        return out.tobytes()

    def histogram(self) -> np.ndarray:
        h = np.zeros(_abi.HIST_BINS, dtype=np.uint64)
        self.ctx._ck(self.lib.mfkc_kset_histogram(self.h, h.ctypes.data_as(_abi.u64p)))
        return h

    def sequences(self, freq_threshold: int, len_threshold: int):
        """seq-builder (src/algo/SequencesFinders.java:13-31): [(sequence, av_weight, min_weight, max_weight)]"""
        ns, nb = C.c_uint64(), C.c_uint64()
        self.ctx._ck(self.lib.mfkc_kset_sequences_begin(self.h, freq_threshold, len_threshold, C.byref(ns), C.byref(nb)))
        n = ns.value
        off = np.zeros(n + 1, dtype=np.uint64)
        bases = np.zeros(max(nb.value, 1), dtype=np.uint8)
        av, lo, hi = (np.zeros(max(n, 1), dtype=np.uint32) for _ in range(3))
        self.ctx._ck(self.lib.mfkc_kset_sequences_fetch(self.h, off.ctypes.data_as(_abi.u64p), _ptr(bases), _ptr(av), _ptr(lo), _ptr(hi)))
        text = bases.tobytes().decode("latin-1")
        return [(text[int(off[i]):int(off[i + 1])], int(av[i]), int(lo[i]), int(hi[i])) for i in range(n)]


    def components(self, min_component_size: int = 1000, max_component_size: int = 10000):
        """component-cutter's graph half (src/algo/ComponentsBuilder.java:24-31): [(weight, [k-mers ascending], usedFreqThreshold)]
        in ConnectedComponent.compareTo order"""
        nc, nk = C.c_uint64(), C.c_uint64()
        self.ctx._ck(self.lib.mfkc_kset_components_begin(self.h, min_component_size, max_component_size, C.byref(nc), C.byref(nk)))
        n = nc.value
        off = np.zeros(n + 1, dtype=np.uint64)
        keys = np.zeros(max(nk.value, 1), dtype=np.int64)
        weight = np.zeros(max(n, 1), dtype=np.int64)
        thr = np.zeros(max(n, 1), dtype=np.int32)
        self.ctx._ck(self.lib.mfkc_kset_components_fetch(self.h, off.ctypes.data_as(_abi.u64p), _ptr(keys), _ptr(weight), _ptr(thr)))
        return [(int(weight[i]), [int(x) for x in keys[int(off[i]):int(off[i + 1])]], int(thr[i])) for i in range(n)]


def kmers_filter(ctx: "_Ctx", inputs: Sequence[bytes], filters: Sequence[bytes], b: int = 1, max_thresh: int = 0):
    """kmers-filter (src/tools/KmersFilter.java:94-110): per input file -> (hm.size(), filtered records)"""
    out = []
    with KmerSet.load(ctx, filters, b) as flt:
        for data in inputs:
            with KmerSet.load(ctx, [data], b) as hm:
                out.append((hm.size(), hm.select(flt, b, max_thresh * len(filters))))
    return out


def unique_kmers_multi(ctx: "_Ctx", inputs: Sequence[bytes], filters: Sequence[bytes], b: int = 1, min_samples: int = 1,
                       max_samples: int = 1):
    """unique-kmers-multi (src/tools/UniqueKmersMultipleSamplesFinder.java:97-148) -> (hm.size(), {i: records})"""
    with KmerSet(ctx) as hm, KmerSet(ctx) as cnt:
        for data in inputs:
            with KmerSet.load(ctx, [data], b) as tmp:
                hm.update(tmp, KmerSet.ADD, b)
                cnt.update(tmp, KmerSet.INC, b)
        for data in filters:
            with KmerSet.load(ctx, [data], b) as flt:
                hm.update(flt, KmerSet.ZERO, b)
        return hm.size(), {i: hm.select(cnt, b, i - 1) for i in range(min_samples, max_samples + 1)}


def kmers_samples_counter(ctx: "_Ctx", inputs: Sequence[bytes], b: int = 1):
    """kmers-samples-counter (src/tools/KmersSamplesCounter.java:90-119) -> (hm.size(), records, histogram)"""
    with KmerSet.load(ctx, inputs, b) as hm:
        hm.reset_values()
        for data in inputs:
            with KmerSet.load(ctx, [data], b) as one:
                hm.update(one, KmerSet.INC, b)
        return hm.size(), hm.select(None, 0), hm.histogram()


class FeaturesCalculator(_Ctx):
    """features-calculator on the device (K6/K7/K8)."""

    def load_components(self, comps: Sequence[Sequence[int]]):
        off = np.zeros(len(comps) + 1, dtype=np.uint64)
        if comps:
            off[1:] = np.cumsum([len(c) for c in comps], dtype=np.uint64)
        flat = np.array([k for c in comps for k in c], dtype=np.uint64).view(np.int64)
        if flat.size == 0:
            flat = np.zeros(1, dtype=np.int64)
        self.n_comp = len(comps)
        self._ck(self.lib.mfkc_fc_load_components(self.h, _ptr(flat), _ptr(off), len(comps)))

    def set_selected(self, records: Optional[bytes]):
        if records is None:
            self._ck(self.lib.mfkc_fc_set_selected(self.h, None, 0))
            return
        a = np.frombuffer(records, dtype=np.uint8) if records else np.zeros(1, dtype=np.uint8)
        self._ck(self.lib.mfkc_fc_set_selected(self.h, _ptr(a), len(records) // 10))

    def reset_values(self):
        self._ck(self.lib.mfkc_fc_reset_values(self.h))

    def add_records(self, records: bytes, chunk: int = 16777200):
        a = np.frombuffer(records, dtype=np.uint8)
        for s in range(0, len(records), chunk):
            part = a[s:s + chunk]
            self._ck(self.lib.mfkc_fc_add_records(self.h, _ptr(part), part.nbytes // 10))

    def add_reads(self, reads: Sequence[str]):
        b, o = pack_reads(reads)
        self._ck(self.lib.mfkc_fc_add_reads(self.h, _ptr(b), _ptr(o), len(reads)))

    def add_emitted(self, counter: "KmerCounter"):
        """the records `counter` (same GPU) selected with its last emit_begin, straight from device memory"""
        self._ck(self.lib.mfkc_fc_add_emitted(self.h, counter.h))

    def features(self, threshold: int = 0):
        n = self.n_comp
        vec = np.zeros(max(n, 1), dtype=np.int64)
        found = np.zeros(max(n, 1), dtype=np.uint64)
        cnt = np.zeros(max(n, 1), dtype=np.uint64)
        self._ck(self.lib.mfkc_fc_features(self.h, threshold, _ptr(vec), _ptr(found), _ptr(cnt)))
        return vec[:n], found[:n], cnt[:n]


# ---- host-side pieces (CPU, no GPU needed) ------------------------------------------------
class ReaderError(_abi.MfkcError):
    pass


def read_file(path: str, batch_reads: int = 1 << 15, batch_bases: int = 1 << 24) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
    """Yield (bases, offsets) batches of the reads the reference's parser would keep."""
    lib = _abi.load()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib.mfkc_reader_open(path.encode(), C.byref(h), err, 512)
    if rc != 0:
        raise ReaderError(rc, err.value.decode())
    try:
        bases = np.empty(batch_bases, dtype=np.uint8)
        offsets = np.empty(batch_reads + 1, dtype=np.uint64)
        n = C.c_uint32()
        while True:
            rc = lib.mfkc_reader_next(h, _ptr(bases), bases.nbytes, _ptr(offsets), batch_reads, C.byref(n))
            if rc == _abi.E_BADARG:                       # one read longer than the buffer: it stays pending, take it with a larger one
                pending = C.c_uint64()
                lib.mfkc_reader_pending_bases(h, C.byref(pending))
                if pending.value > bases.nbytes:
                    bases = np.empty(pending.value + pending.value // 8, dtype=np.uint8)
                    rc = lib.mfkc_reader_next(h, _ptr(bases), bases.nbytes, _ptr(offsets), batch_reads, C.byref(n))
            if rc != 0:
                raise ReaderError(rc, lib.mfkc_reader_error(h).decode())
            if n.value == 0:
                break
            nb = int(offsets[n.value])
            yield bases[:max(nb, 1)].copy(), offsets[: n.value + 1].copy()
    finally:
        lib.mfkc_reader_close(h)


def read_file_reads(path: str) -> List[str]:
    out: List[str] = []
    for b, o in read_file(path):
        s = b.tobytes().decode("latin-1")
        out += [s[int(o[i]):int(o[i + 1])] for i in range(len(o) - 1)]
    return out


def reader_name(path: str) -> str:
    lib = _abi.load()
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib.mfkc_reader_open(path.encode(), C.byref(h), err, 512)
    if rc != 0:
        raise ReaderError(rc, err.value.decode())
    try:
        return lib.mfkc_reader_name(h).decode()
    finally:
        lib.mfkc_reader_close(h)


def synth_cfg(**kw) -> _abi.SynthCfg:
    lib = _abi.load()
    c = _abi.SynthCfg()
    lib.mfkc_synth_defaults(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def synth_reads_host(cfg: _abi.SynthCfg, first: int, n: int) -> np.ndarray:
    lib = _abi.load()
    out = np.empty((n, cfg.read_len), dtype=np.uint8)
    rc = lib.mfkc_synth_reads_host(C.byref(cfg), first, n, _ptr(out))
    if rc != 0:
        raise _abi.MfkcError(rc, "mfkc_synth_reads_host")
    return out


def write_stat_file(path: str, hist: np.ndarray):
    lib = _abi.load()
    rc = lib.mfkc_write_stat_file(path.encode(), hist.ctypes.data_as(_abi.u64p))
    if rc != 0:
        raise _abi.MfkcError(rc, "cannot write " + path)
