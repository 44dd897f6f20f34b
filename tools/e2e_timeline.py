#!/usr/bin/env python
"""Host-side timeline of one end-to-end step (cfg2) through the C ABI: where the wall time goes."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metafast_b200 as m

N, L, K, B = int(os.environ.get("MFKC_BENCH_READS", 20_000_000)), 150, 31, 1_000_000
kc = m.KmerCounter(K, device=0, expected_kmers=N * (L - K + 1))
cfg = m.synth_cfg()
d_b = kc.device_alloc(N * L); d_o = kc.device_alloc((N + 1) * 8)
kept = C.c_uint64()
kc._ck(kc.lib.mfkc_synth_reads_device(kc.h, C.byref(cfg), 0, N, C.c_void_p(d_b), C.c_void_p(d_o), C.byref(kept)))
n = kept.value
hb = kc.pinned(n * L); ho = kc.pinned((n + 1) * 8, np.uint64)
kc.d2h(hb, d_b); kc.d2h(ho, d_o)
out = kc.pinned(200_000_000 * 10)
for it in range(4):
    kc.sync(); t = [time.perf_counter()]
    kc.reset(); t.append(time.perf_counter())
    for s in range(0, n, B):
        e = min(n, s + B); kc.submit(hb, ho[s:e + 1])
    t.append(time.perf_counter())
    kc.flush(); t.append(time.perf_counter())
    ng = kc.emit_begin(2); t.append(time.perf_counter())
    w = C.c_size_t(); pos = 0
    while pos < ng * 10:
        kc._ck(kc.lib.mfkc_emit_next(kc.h, C.c_void_p(out.ctypes.data + pos), out.nbytes - pos, C.byref(w)))
        if not w.value: break
        pos += w.value
    t.append(time.perf_counter())
    kc.histogram(); t.append(time.perf_counter())
    names = ["reset", "submit x%d" % ((n + B - 1) // B), "flush", "emit_begin", "emit_next", "histogram"]
    print("iter %d total %.1f ms | " % (it, 1e3 * (t[-1] - t[0])) + "  ".join("%s %.1f" % (a, 1e3 * (t[i + 1] - t[i])) for i, a in enumerate(names)), flush=True)
