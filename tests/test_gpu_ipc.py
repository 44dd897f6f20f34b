"""The peer-memory exchange between PROCESSES: two ranks on one GPU, staging buffers exported with
mfkc_p2p_export and opened with mfkc_p2p_attach (real CUDA IPC handles), bin-local count (default) and the
region-blocked table flavour.  Plain multiprocessing is the plumbing (pipes carry the 128-byte handles, the
per-owner totals and the results); the merged result must equal the oracle's single count."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank_main(rank, world, flavour, conn, n, k):
    try:
        sys.path.insert(0, ROOT)
        import metafast_b200 as m
        cfg = m.synth_cfg(total_genome_bp=120000, n_genomes=4, n_read_ppm=0)
        raw = m.synth_reads_host(cfg, 0, n)
        bases = np.ascontiguousarray(raw).reshape(-1)
        offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(cfg.read_len)
        variant = m.VARIANT_HASH if flavour == "bins" else m.VARIANT_HASH_TABLE
        kc = m.KmerCounter(k, n_shards=world, shard_id=rank, variant=variant)
        kmers = n * (cfg.read_len - k + 1)
        if flavour == "bins":
            kc.p2p_stage_create_bins(24, int(kmers * 0.19 / (world * world * 24) * 2) + 16, 1 << 16)
        else:
            kc.p2p_stage_create(3, 2 * kmers // 5 // (world << 3) + 64)
        conn.send(kc.p2p_export())
        handles = conn.recv()                                  # every rank's 128 bytes, relayed by the parent
        for r in range(world):
            kc.p2p_attach(r, None if r == rank else handles[r])
        for rep in range(2):
            kc.reset()
            conn.send("ready"); conn.recv()                    # barrier: nobody clears a buffer a peer still reads
            kc.p2p_stage_reset()
            lo, hi = rank * n // world, (rank + 1) * n // world
            kc.p2p_submit(bases, offsets[lo: hi + 1])
            conn.send(kc.p2p_counts(world))                    # synchronises this rank's extraction
            totals = conn.recv()
            kc.p2p_drain(sum(t[rank] for t in totals))
            kc.flush()
            rec = kc.emit(1)
            st = kc.stats()
            conn.send((rec, kc.histogram(), st))
            conn.recv()                                        # barrier: every rank has fetched its results
        kc.close()
        conn.send("done")
    except Exception as ex:                                    # surface the failure in the parent instead of hanging it
        import traceback
        conn.send(("error", traceback.format_exc()))


def _recv(conn, timeout=120):
    if not conn.poll(timeout):
        raise TimeoutError("rank did not answer")
    v = conn.recv()
    if isinstance(v, tuple) and len(v) == 2 and v[0] == "error":
        raise RuntimeError(v[1])
    return v


@pytest.mark.parametrize("flavour", ["bins", "table"])
def test_two_processes_cuda_ipc(built, flavour):
    import metafast_b200 as m
    from metafast_b200.sharded import merge_sorted_records
    from tests import _oracle_c
    world, n, k = 2, 8000, 31
    cfg = m.synth_cfg(total_genome_bp=120000, n_genomes=4, n_read_ppm=0)
    raw = m.synth_reads_host(cfg, 0, n)
    bases = np.ascontiguousarray(raw).reshape(-1)
    offsets = np.arange(n + 1, dtype=np.uint64) * np.uint64(cfg.read_len)
    want_rec, want_hist, want_distinct, want_stats = _oracle_c.count(bases, offsets, k, 1, P=2)
    ctx = mp.get_context("spawn")
    pipes, procs = [], []
    try:
        for r in range(world):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_rank_main, args=(r, world, flavour, b, n, k), daemon=True)
            p.start()
            pipes.append(a); procs.append(p)
        handles = [_recv(c) for c in pipes]
        for c in pipes:
            c.send(handles)
        for rep in range(2):
            for c in pipes:
                assert _recv(c) == "ready"
            for c in pipes:
                c.send("go")
            totals = [_recv(c) for c in pipes]
            for c in pipes:
                c.send(totals)
            res = [_recv(c) for c in pipes]
            for c in pipes:
                c.send("ok")
            assert merge_sorted_records([r_[0] for r_ in res]) == want_rec
            assert (sum(r_[1] for r_ in res) == want_hist).all()
            assert sum(r_[2]["distinct"] for r_ in res) == want_distinct
            assert [sum(r_[2][f] for r_ in res) for f in ("total_seq", "good_seq", "total_len", "good_len")] == want_stats
        for c in pipes:
            assert _recv(c) == "done"
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
