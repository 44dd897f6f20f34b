import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import metafast_b200 as m
from oracle import oracle as orc
from tests import _oracle_c
for n_shards in (3, 3, 8):
    cfg = m.synth_cfg(total_genome_bp=100000, n_genomes=4, n_read_ppm=0)
    n = 4000
    raw = m.synth_reads_host(cfg, 0, n)
    bases = np.ascontiguousarray(raw).reshape(-1)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(cfg.read_len)
    want_rec, want_hist, _, _ = _oracle_c.count(bases, offsets, 31, 1, P=2)
    want = dict(orc.load_kmers_bin(want_rec))
    lib = m.load()
    shards = [m.KmerCounter(31, n_shards=n_shards, shard_id=s) for s in range(n_shards)]
    src = shards[0]
    d_b = src.device_alloc(bases.nbytes); d_o = src.device_alloc(offsets.nbytes)
    src.h2d(d_b, bases); src.h2d(d_o, offsets)
    cap = bases.size
    d_keys = src.device_alloc(cap * 8)
    counts = src.extract_bucketed(d_b, d_o, n, bases.size, d_keys, cap, n_shards)
    keys = np.empty(sum(counts), dtype=np.uint64)
    src.d2h(keys, d_keys)
    truth_keys = orc.canonical_kmers_np([bytes(r).decode() for r in raw], 31)
    print("n_shards", n_shards, "counts", counts, "keys multiset equal:", bool((np.sort(keys) == np.sort(truth_keys)).all()))
    pos = 0
    got = {}
    for s in range(n_shards):
        part = keys[pos:pos + counts[s]]; pos += counts[s]
        d_part = shards[s].device_alloc(max(part.nbytes, 8))
        shards[s].h2d(d_part, part)
        shards[s].count_keys_device(d_part, part.size)
        shards[s].flush()
        rec = shards[s].emit(1)
        uniq, cnt = np.unique(part, return_counts=True)
        exp = {int(k): min(int(c), 32767) for k, c in zip(uniq, cnt) if c > 1}
        g = dict(orc.load_kmers_bin(rec))
        bad = {k: (g.get(k), exp.get(k)) for k in set(g) | set(exp) if g.get(k) != exp.get(k)}
        print(" shard", s, "part", part.size, "distinct", len(uniq), "stats", shards[s].stats()["distinct"], "records", len(g), "expected", len(exp), "mismatches", len(bad), list(bad.items())[:5], shards[s].bin_stats())
        got.update(g)
    print(" merged equal:", got == want)
    for s in shards:
        s.close()
