import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metafast_b200 as m
with m.KmerCounter(31) as kc:
    nupd = 1 << 30
    for win_mb in (4, 8, 16):
        for bpw in (296, 592):
            for mode in (3, 4, 5):
                ms = kc.gups(8 << 30, nupd, mode, win_mb << 20, bpw)
                print("8GiB windowed win=%dMiB bpw=%d mode=%d: %.2f ms %.2f Gupd/s" % (win_mb, bpw, mode, ms, nupd / ms / 1e6), flush=True)
