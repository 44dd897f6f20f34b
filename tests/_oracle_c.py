"""ctypes access to oracle/_build/liboracle.so (the C restatement of the reference's CPU
algorithm).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
BIN = os.path.join(ROOT, "oracle", "_build", "ref_cpu")
_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB)
        lib.orc_map_new.restype = C.c_void_p
        lib.orc_map_new.argtypes = [C.c_int]
        lib.orc_map_free.argtypes = [C.c_void_p]
        lib.orc_map_size.restype = C.c_uint64
        lib.orc_map_size.argtypes = [C.c_void_p]
        lib.orc_map_read_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        lib.orc_count_reads.restype = C.c_int
        lib.orc_count_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int]
        lib.orc_emit.restype = C.c_uint64
        lib.orc_emit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        lib.orc_parse_file.restype = C.c_void_p
        lib.orc_parse_file.argtypes = [C.c_char_p]
        lib.orc_reads_free.argtypes = [C.c_void_p]
        for f, t in (("orc_reads_bases", C.c_void_p), ("orc_reads_offsets", C.c_void_p),
                     ("orc_reads_count", C.c_uint64), ("orc_reads_nbases", C.c_uint64),
                     ("orc_reads_error", C.c_char_p)):
            getattr(lib, f).restype = t
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.orc_features_kmers.restype = C.c_int
        lib.orc_features_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                           C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def count(bases: np.ndarray, offsets: np.ndarray, k: int, threshold: int, min_len: int = 0, P: int = 4,
          sort: bool = True):
    """-> (records bytes, hist uint64[32768], distinct, read stats[4])"""
    lib = load()
    hm = lib.orc_map_new(P)
    try:
        rc = lib.orc_count_reads(hm, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, k, min_len, P)
        if rc != 0:
            raise RuntimeError("orc_count_reads rc=%d" % rc)
        distinct = lib.orc_map_size(hm)
        hist = np.zeros(32768, dtype=np.uint64)
        good = lib.orc_emit(hm, threshold, None, 0, None, 0)
        out = np.zeros(max(good * 10, 1), dtype=np.uint8)
        lib.orc_emit(hm, threshold, out.ctypes.data, good, hist.ctypes.data, 1 if sort else 0)
        st = (C.c_uint64 * 4)()
        lib.orc_map_read_stats(hm, st)
        return out[: good * 10].tobytes(), hist, distinct, list(st)
    finally:
        lib.orc_map_free(hm)


def parse_file(path: str):
    lib = load()
    r = lib.orc_parse_file(path.encode())
    try:
        err = lib.orc_reads_error(r)
        if err:
            raise RuntimeError(err.decode())
        n = lib.orc_reads_count(r)
        nb = lib.orc_reads_nbases(r)
        bases = np.ctypeslib.as_array(C.cast(lib.orc_reads_bases(r), C.POINTER(C.c_uint8)), shape=(max(nb, 1),)).copy() \
            if nb else np.zeros(1, dtype=np.uint8)
        offsets = np.ctypeslib.as_array(C.cast(lib.orc_reads_offsets(r), C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
        return bases, offsets
    finally:
        lib.orc_reads_free(r)


def features_kmers(comps, records: bytes, selected: bytes = None, threshold: int = 0):
    lib = load()
    off = np.zeros(len(comps) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(c) for c in comps], dtype=np.uint64)
    flat = np.array([k for c in comps for k in c], dtype=np.uint64)
    if flat.size == 0:
        flat = np.zeros(1, dtype=np.uint64)
    rec = np.frombuffer(records, dtype=np.uint8) if records else np.zeros(1, dtype=np.uint8)
    n = len(comps)
    vec = np.zeros(max(n, 1), dtype=np.int64)
    found = np.zeros(max(n, 1), dtype=np.uint64)
    cnt = np.zeros(max(n, 1), dtype=np.uint64)
    if selected is not None:
        sel = np.frombuffer(selected, dtype=np.uint8) if selected else np.zeros(1, dtype=np.uint8)
        sp, sn = sel.ctypes.data, len(selected) // 10
    else:
        sp, sn = None, 0
    lib.orc_features_kmers(flat.ctypes.data, off.ctypes.data, n, rec.ctypes.data, len(records) // 10, sp, sn,
                           threshold, vec.ctypes.data, found.ctypes.data, cnt.ctypes.data)
    return vec[:n], found[:n], cnt[:n]
