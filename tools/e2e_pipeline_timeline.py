"""Timeline of the pipelined end-to-end path (two contexts, host buffers): where a sample's wall time goes.
   python tools/e2e_pipeline_timeline.py [n_steps]"""
import ctypes as C, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metafast_b200 as m

K, B, L = 31, 2, 150
N = int(os.environ.get("MFKC_BENCH_READS", 20_000_000)); BATCH = int(os.environ.get("MFKC_BENCH_BATCH", 1_000_000))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
LANES = int(os.environ.get('TL_LANES', 3))
kcs = [m.KmerCounter(K, expected_kmers=N * (L - K + 1)) for _ in range(LANES)]
kc = kcs[0]
d_b = kc.device_alloc(N * L); d_o = kc.device_alloc((N + 1) * 8)
kept = C.c_uint64(); cfg = m.synth_cfg()
kc._ck(kc.lib.mfkc_synth_reads_device(kc.h, C.byref(cfg), 0, N, C.c_void_p(d_b), C.c_void_p(d_o), C.byref(kept)))
n = kept.value
h_b = kc.pinned(n * L); h_o = kc.pinned((n + 1) * 8, np.uint64)
kc.d2h(h_b, d_b); kc.d2h(h_o, d_o)
kc.device_free(d_b); kc.device_free(d_o)
outs = [c.pinned(130_000_000 * 10) for c in kcs]
dummy = kc.device_alloc(1_200_000_000)
link = threading.Lock()
T0 = time.perf_counter()
log = []

def sample(j, i, use_lock=True):
    c = kcs[j]
    t = [time.perf_counter()]
    if use_lock:
        link.acquire()
    t.append(time.perf_counter())
    c.reset()
    for s in range(0, n, BATCH):
        e = min(n, s + BATCH)
        c.submit(h_b, h_o[s:e + 1])
    t.append(time.perf_counter())
    if use_lock:
        link.release()
    c.flush(); t.append(time.perf_counter())
    if os.environ.get("TL_SKIP_COUNT"):
        ng = 120_000_000
        c.lib.mfkc_emit_begin  # (not called: only the copy back of stale records)
    else:
        ng = c.emit_begin(B)
    t.append(time.perf_counter())
    w = C.c_size_t(); pos = 0
    while pos < ng * 10 and not os.environ.get("TL_SKIP_D2H") and not os.environ.get("TL_SKIP_COUNT"):
        c._ck(c.lib.mfkc_emit_next(c.h, C.c_void_p(outs[j].ctypes.data + pos), outs[j].nbytes - pos, C.byref(w)))
        if not w.value:
            break
        pos += w.value
    if os.environ.get("TL_SKIP_COUNT"):        # a plain device-to-host copy of the same size instead
        c.d2h(outs[j][:1_200_000_000], dummy)
    t.append(time.perf_counter())
    if not os.environ.get("TL_SKIP_COUNT"):
        c.histogram()
    t.append(time.perf_counter())
    log.append((i, j, [1e3 * (x - T0) for x in t]))

def run(nsteps, use_lock=True, lanes=2):
    def lane(j):
        for i in range(j, nsteps, lanes):
            sample(j, i, use_lock)
    ts = [threading.Thread(target=lane, args=(j,)) for j in range(lanes)]
    t0 = time.perf_counter()
    [t.start() for t in ts]; [t.join() for t in ts]
    return 1e3 * (time.perf_counter() - t0) / nsteps

run(2 * LANES, lanes=LANES)
for name, kw in (("serial (1 context)", dict(lanes=1)), ("pipelined, link lock", dict(lanes=LANES))):
    del log[:]
    ms = run(steps, **kw)
    print("%s: %.1f ms per sample" % (name, ms))
    for i, j, t in sorted(log):
        print("  step %d ctx %d: wait %.0f | submit %.0f-%.0f (%.0f) | flush +%.0f | count+sort +%.0f | D2H +%.0f | hist +%.0f" %
              (i, j, t[1] - t[0], t[1], t[2], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5]))
