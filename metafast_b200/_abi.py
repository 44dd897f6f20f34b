"""ctypes binding of libmfkc.so (include/mfkc.h).

This is the only way Python code in this repository reaches the product: plain
pointers and sizes through the C ABI, exactly what a Java host would bind with
Panama FFM.  There is no fallback: if the shared library is missing, or no CUDA
device is present when a context is created, the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MFKC_LIBRARY") or os.path.join(_HERE, "lib", "libmfkc.so")     # MFKC_LIBRARY: A/B builds of the same ABI

MFKC_OK = 0
E_BADARG, E_CUDA, E_NCCL, E_TABLE_FULL, E_OOM, E_STATE, E_IO, E_FORMAT = -1, -2, -3, -4, -5, -6, -7, -8
VARIANT_HASH, VARIANT_SORT, VARIANT_HASH_DIRECT, VARIANT_HASH_TABLE = 0, 1, 2, 3
MAX_COUNT = 32767
HIST_BINS = 32768

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)


class MfkcCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("k", C.c_int32),
        ("min_seq_len", C.c_int32),
        ("device", C.c_int32),
        ("variant", C.c_int32),
        ("n_shards", C.c_int32),
        ("shard_id", C.c_int32),
        ("reserved0", C.c_int32),
        ("table_slots", C.c_uint64),
        ("expected_distinct", C.c_uint64),
        ("max_table_bytes", C.c_uint64),
        ("staging_bytes", C.c_uint64),
        ("region_shift", C.c_uint32),
        ("reserved2", C.c_uint32),
        ("expected_kmers", C.c_uint64),
        ("reserved1", C.c_uint64 * 1),
    ]


class SynthCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("n_genomes", C.c_uint32),
        ("seed", C.c_uint64),
        ("total_genome_bp", C.c_uint64),
        ("read_len", C.c_uint32),
        ("sample", C.c_uint32),
        ("err_ppm_first", C.c_uint32),
        ("err_ppm_last", C.c_uint32),
        ("n_read_ppm", C.c_uint32),
        ("poly_tail_ppm", C.c_uint32),
        ("reserved", C.c_uint64 * 4),
    ]


# name -> (restype, argtypes): every symbol include/mfkc.h declares
SIGNATURES = {
    "mfkc_abi_version": (C.c_int, []),
    "mfkc_device_count": (C.c_int, []),
    "mfkc_create": (C.c_int, [C.POINTER(MfkcCfg), C.POINTER(C.c_void_p)]),
    "mfkc_destroy": (None, [C.c_void_p]),
    "mfkc_last_error": (C.c_char_p, [C.c_void_p]),
    "mfkc_reset": (C.c_int, [C.c_void_p]),
    "mfkc_pinned_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "mfkc_pinned_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfkc_submit_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "mfkc_submit_reads_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]),
    "mfkc_flush": (C.c_int, [C.c_void_p]),
    "mfkc_stats": (C.c_int, [C.c_void_p, u64p]),
    "mfkc_library_name": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t]),
    "mfkc_bin_stats": (C.c_int, [C.c_void_p, u64p]),
    "mfkc_merge_records": (C.c_int, [C.c_void_p, u64p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]),
    "mfkc_histogram": (C.c_int, [C.c_void_p, u64p]),
    "mfkc_emit_begin": (C.c_int, [C.c_void_p, C.c_int32, u64p]),
    "mfkc_emit_next": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "mfkc_emit_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), u64p]),
    "mfkc_extract_bucketed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64,
                                        C.c_void_p, C.c_uint64, u64p]),
    "mfkc_count_keys_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "mfkc_owner_shard": (C.c_uint32, [C.c_uint64, C.c_uint32]),
    "mfkc_skm_extract_bucketed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64,
                                            C.c_void_p, C.c_uint64, u64p, u64p]),
    "mfkc_skm_count_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "mfkc_skm_count_wait": (C.c_int, [C.c_void_p]),
    "mfkc_p2p_stage_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64]),
    "mfkc_p2p_stage_create_bins": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64]),
    "mfkc_p2p_bin_geometry": (C.c_int, [C.c_uint64, C.c_uint32, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_uint32), u64p, u64p]),
    "mfkc_p2p_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfkc_p2p_attach": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "mfkc_p2p_attach_ctx": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "mfkc_p2p_stage_reset": (C.c_int, [C.c_void_p]),
    "mfkc_p2p_extract": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]),
    "mfkc_p2p_submit_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "mfkc_p2p_counts": (C.c_int, [C.c_void_p, u64p]),
    "mfkc_p2p_drain": (C.c_int, [C.c_void_p, C.c_uint64]),
    "mfkc_fc_load_components": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "mfkc_fc_set_selected": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "mfkc_fc_reset_values": (C.c_int, [C.c_void_p]),
    "mfkc_fc_add_records": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "mfkc_fc_add_emitted": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfkc_fc_add_reads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]),
    "mfkc_fc_features": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mfkc_kset_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mfkc_kset_destroy": (None, [C.c_void_p]),
    "mfkc_kset_load_records": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32]),
    "mfkc_kset_load_finish": (C.c_int, [C.c_void_p]),
    "mfkc_kset_size": (C.c_int, [C.c_void_p, u64p]),
    "mfkc_kset_reset_values": (C.c_int, [C.c_void_p]),
    "mfkc_kset_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int32]),
    "mfkc_kset_select_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, u64p]),
    "mfkc_kset_select_next": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "mfkc_kset_histogram": (C.c_int, [C.c_void_p, u64p]),
    "mfkc_kset_sequences_begin": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, u64p, u64p]),
    "mfkc_kset_sequences_fetch": (C.c_int, [C.c_void_p, u64p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mfkc_kset_components_begin": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, u64p, u64p]),
    "mfkc_kset_components_fetch": (C.c_int, [C.c_void_p, u64p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mfkc_reader_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]),
    "mfkc_reader_next": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "mfkc_reader_pending_bases": (C.c_int, [C.c_void_p, u64p]),
    "mfkc_reader_counters": (C.c_int, [C.c_void_p, u64p]),
    "mfkc_reader_error": (C.c_char_p, [C.c_void_p]),
    "mfkc_reader_name": (C.c_char_p, [C.c_void_p]),
    "mfkc_reader_close": (None, [C.c_void_p]),
    "mfkc_write_stat_file": (C.c_int, [C.c_char_p, u64p]),
    "mfkc_synth_defaults": (None, [C.POINTER(SynthCfg)]),
    "mfkc_synth_reads_host": (C.c_int, [C.POINTER(SynthCfg), C.c_uint64, C.c_uint64, C.c_void_p]),
    "mfkc_synth_reads_device": (C.c_int, [C.c_void_p, C.POINTER(SynthCfg), C.c_uint64, C.c_uint64, C.c_void_p,
                                          C.c_void_p, u64p]),
    "mfkc_device_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "mfkc_device_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mfkc_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "mfkc_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "mfkc_device_sync": (C.c_int, [C.c_void_p]),
    "mfkc_timer_start": (C.c_int, [C.c_void_p]),
    "mfkc_timer_stop_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "mfkc_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "mfkc_profile_reset": (C.c_int, [C.c_void_p]),
    "mfkc_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), u64p]),
    "mfkc_profile_name": (C.c_char_p, [C.c_int]),
    "mfkc_gups": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_float)]),
    "mfkc_gups_ex": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint32,
                               C.POINTER(C.c_float)]),
}

_lib = None


class MfkcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libmfkc error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


def load() -> C.CDLL:
    """Load libmfkc.so and bind every declared symbol.  Raises if the library has
    not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it first (__graft_entry__.build() or `make -C metafast_b200`). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError = symbol missing from the library
        fn.restype = res
        fn.argtypes = args
    if lib.mfkc_abi_version() != 1:
        raise ImportError("libmfkc ABI version mismatch")
    _lib = lib
    return lib
