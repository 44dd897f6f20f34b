"""metafast_b200 -- B200-native (sm_100a) k-mer counting hot path of MetaFast.

Only the hot path lives here: CUDA kernels + C ABI (``csrc/``, built into
``lib/libmfkc.so``), the C++ host tool (``bin/mfkc_cli``) and a thin ctypes layer.
"""
from ._abi import MfkcError, load, LIB_PATH, VARIANT_HASH, VARIANT_SORT, VARIANT_HASH_DIRECT, VARIANT_HASH_TABLE, MAX_COUNT, HIST_BINS  # noqa: F401
from .counter import (KmerCounter, FeaturesCalculator, KmerSet, kmers_filter, unique_kmers_multi, kmers_samples_counter, pack_reads, read_file, read_file_reads,  # noqa: F401
                      reader_name, synth_cfg, synth_reads_host, write_stat_file)

__version__ = "0.1.0"
