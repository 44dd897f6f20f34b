"""Quick GPU probe: GUPS random-sector peak + a reduced config-2 run with per-kernel timings."""
import ctypes as C, sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metafast_b200 as m

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
variant = {"sort": m.VARIANT_SORT, "direct": m.VARIANT_HASH_DIRECT}.get(sys.argv[2] if len(sys.argv) > 2 else "", m.VARIANT_HASH)
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
hint = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cfg = m.synth_cfg()
KK = int(os.environ.get('MFKC_K', 31))
stg = int(sys.argv[5]) if len(sys.argv) > 5 else 0
slots = int(sys.argv[6]) if len(sys.argv) > 6 else 0
with m.KmerCounter(KK, variant=variant, expected_distinct=hint, staging_bytes=stg, table_slots=slots) as kc:
    for gb in ():
        for dep in (False, True):
            nupd = 1 << 29
            ms = kc.gups(gb << 30, nupd, 1 if dep else 0)
            print("gups table=%dGiB dependent=%d: %.2f ms  %.2f Gupd/s  %.1f GB/s (64B/upd)" % (gb, dep, ms, nupd / ms / 1e6, nupd * 64 / ms / 1e6), flush=True)
    # device-resident synthetic reads
    L = cfg.read_len
    d_b = kc.device_alloc(n_reads * L); d_o = kc.device_alloc((n_reads + 1) * 8)
    kept = C.c_uint64()
    t = time.time()
    kc._ck(kc.lib.mfkc_synth_reads_device(kc.h, C.byref(cfg), 0, n_reads, C.c_void_p(d_b), C.c_void_p(d_o), C.byref(kept)))
    print("synth %d reads (%d kept) in %.2fs" % (n_reads, kept.value, time.time() - t), flush=True)
    nk = kept.value
    for rep in range(3):
        kc.reset()
        kc.profile(enable=True, reset=True)
        kc.timer_start()
        t = time.time()
        for s in range(0, nk, batch):
            e = min(nk, s + batch)
            kc.submit_device(d_b + s * L, d_o + s * 8, e - s, (e - s) * L)
        kc.flush()
        ms_count = kc.timer_stop_ms()
        kc.timer_start()
        ng = kc.emit_begin(2)
        ms_emit = kc.timer_stop_ms()
        st = kc.stats()
        kmers = st["kmers"]
        print("rep %d: count %.2f ms, emit %.2f ms, wall %.3fs, kmers %d distinct %d good %d -> %.2f Gkmer/s (count), %.2f (count+emit); hash-algorithmic %.1f GB/s" % (
            rep, ms_count, ms_emit, time.time() - t, kmers, st["distinct"], ng, kmers / ms_count / 1e6, kmers / (ms_count + ms_emit) / 1e6, kmers * 64.3125 / ms_count / 1e6), flush=True)
        print("   profile:", {k: (round(v[0], 3), v[1]) for k, v in kc.profile().items() if v[1]}, flush=True)
