import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metafast_b200 as m
with m.KmerCounter(31) as kc:
    nupd = 1 << 30
    for gb, mode in ((8, 0), (8, 1), (8, 2)):
        ms = kc.gups(gb << 30, nupd, mode)
        print("gups table=%dGiB mode=%d: %.2f ms  %.2f Gacc/s" % (gb, mode, ms, nupd / ms / 1e6), flush=True)
    for mb in (16, 32, 64, 96):
        for mode in (0, 1, 2):
            ms = kc.gups(mb << 20, nupd, mode)
            print("gups table=%dMiB (L2-resident) mode=%d: %.2f ms  %.2f Gacc/s" % (mb, mode, ms, nupd / ms / 1e6), flush=True)
    for win_mb in (8, 16, 32, 64):
        for bpw in (74, 148, 296, 592):
            ms = kc.gups(8 << 30, nupd, 3, win_mb << 20, bpw)
            print("gups table=8GiB windowed win=%dMiB blocks/window=%d: %.2f ms %.2f Gupd/s" % (win_mb, bpw, ms, nupd / ms / 1e6), flush=True)
