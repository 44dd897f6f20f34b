#!/usr/bin/env python
"""BASELINE configs[2] in small: kmer-counter-many batch mode over S synthetic samples x 5 M reads (k = 31, -b 2),
then features-calculator of every sample's .kmers.bin records against fixed components (10 000 x 2 000 k-mers).
One GPU = one stream of samples (no exchange; G GPUs would each take every G-th sample).  Prints per-phase times."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metafast_b200 as m

S = int(os.environ.get("CFG3_SAMPLES", 6)); N = int(os.environ.get("CFG3_READS", 5_000_000)); L, K, B = 150, 31, 1_000_000
NC, CS = int(os.environ.get("CFG3_COMPONENTS", 10_000)), int(os.environ.get("CFG3_COMP_SIZE", 2_000))
kc = m.KmerCounter(K, device=0, expected_kmers=N * (L - K + 1))
fc = m.FeaturesCalculator(K, device=0)
d_b = kc.device_alloc(N * L); d_o = kc.device_alloc((N + 1) * 8)
h_b = kc.pinned(N * L); h_o = kc.pinned((N + 1) * 8, np.uint64)
out = kc.pinned(120_000_000 * 10)
comps_loaded = False
tot = {"count+emit": 0.0, "features": 0.0}
for s in range(S):
    cfg = m.synth_cfg(sample=s)
    kept = C.c_uint64()
    kc._ck(kc.lib.mfkc_synth_reads_device(kc.h, C.byref(cfg), 0, N, C.c_void_p(d_b), C.c_void_p(d_o), C.byref(kept)))
    n = kept.value
    kc.d2h(h_b[: n * L], d_b); kc.d2h(h_o[: n + 1], d_o)          # "parsed reads in pinned host memory"
    kc.sync(); t0 = time.perf_counter()
    kc.reset()
    for a in range(0, n, B):
        e = min(n, a + B); kc.submit(h_b, h_o[a:e + 1])
    kc.flush()
    nbytes = kc.emit_into(2, out)
    hist = kc.histogram()
    t1 = time.perf_counter()
    recs = out[:nbytes]
    if not comps_loaded:                                         # fixed components: random good k-mers of sample 0
        keys = recs.reshape(-1, 10)[:, :8].copy().view(">u8").reshape(-1).astype(np.uint64)
        rng = np.random.default_rng(1)
        flat = keys[rng.integers(0, len(keys), NC * CS)].view(np.int64)
        off = (np.arange(NC + 1, dtype=np.uint64) * np.uint64(CS))
        fc._ck(fc.lib.mfkc_fc_load_components(fc.h, flat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), NC))
        fc.n_comp = NC; comps_loaded = True
    fc.sync(); t2 = time.perf_counter()
    fc.reset_values()
    chunk = 16777200
    for a in range(0, nbytes, chunk):
        part = recs[a:a + chunk]
        fc._ck(fc.lib.mfkc_fc_add_records(fc.h, part.ctypes.data_as(C.c_void_p), part.nbytes // 10))
    vec, found, cnt = fc.features(0)
    t3 = time.perf_counter()
    st = kc.stats()
    tot["count+emit"] += t1 - t0; tot["features"] += t3 - t2
    print("sample %d: %d reads, %.0f M k-mers, %.0f M distinct, %.1f M records | count+emit %.1f ms (%.2f Gkmer/s e2e) | features %.1f ms "
          "(%.1f M records/s; vec sum %d, breadth mean %.3f)" % (s, n, st["kmers"] / 1e6, st["distinct"] / 1e6, nbytes / 1e7, 1e3 * (t1 - t0),
          st["kmers"] / (t1 - t0) / 1e9, 1e3 * (t3 - t2), nbytes / 10 / (t3 - t2) / 1e6, int(vec.sum()), float((found / np.maximum(cnt, 1)).mean())), flush=True)
print("per sample: count+emit %.1f ms, features %.1f ms -> 100 samples in %.1f s on one GPU" % (1e3 * tot["count+emit"] / S, 1e3 * tot["features"] / S,
      100 * (tot["count+emit"] + tot["features"]) / S))
