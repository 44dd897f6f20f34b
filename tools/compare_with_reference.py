#!/usr/bin/env python
"""compare_with_reference.py -- for a maintainer who HAS the reference JVM build: check this repository's outputs against
the reference's on the same inputs.

    # reference:  ./metafast.sh -t kmer-counter-many -k 31 -b 1 -i reads/* -w ref_wd
    # this repo:  mfkc_cli     -t kmer-counter-many -k 31 -b 1 -i reads/* -w gpu_wd
    python tools/compare_with_reference.py ref_wd gpu_wd

Compares, for every file name present in both work directories:
  kmers/*.kmers.bin      as MULTISETS of 10-byte records (the reference writes hash-map iteration order, this repository
                         ascending key order; src/io/IOUtils.java:45-71)
  stats/*.stat.txt       byte for byte
  vectors/*.vec|.breadth byte for byte (same components.bin assumed)
  components.bin         as a set of (weight, sorted k-mers) components (order of ties and of the k-mers inside a component is
                         thread / hash-map order in the reference; src/algo/ComponentsBuilder.java)
  matrices/*_original_order.txt  numerically, cell by cell
Exit code 0 = everything present in both trees is identical in that sense.  numpy only; no GPU, no oracle."""
import glob
import os
import struct
import sys

import numpy as np


def sorted_records(path):
    data = np.fromfile(path, dtype=np.uint8)
    if data.size % 10:
        raise SystemExit("%s: size is not a multiple of 10" % path)
    rec = data.reshape(-1, 10)
    return rec[np.lexsort(rec.T[::-1])]


def components(path):
    data = open(path, "rb").read()
    (n,) = struct.unpack_from(">i", data, 0)
    off, out = 4, []
    for _ in range(n):
        size, weight = struct.unpack_from(">iq", data, off)
        off += 12
        keys = np.frombuffer(data, dtype=">i8", count=size, offset=off)
        off += 8 * size
        out.append((weight, tuple(sorted(int(k) for k in keys))))
    return sorted(out)


def matrix(path):
    rows = [ln.rstrip("\n").split("\t") for ln in open(path) if ln.strip()]
    if rows and rows[0][0] == "#":
        names = rows[0][1:]
        return {(r[0], c): float(v) for r in rows[1:] for c, v in zip(names, r[1:])}
    return {(i, j): float(v) for i, r in enumerate(rows) for j, v in enumerate(r)}


def main(ref, new):
    bad = checked = 0

    def pairs(pattern):
        for a in sorted(glob.glob(os.path.join(ref, pattern))):
            b = os.path.join(new, os.path.relpath(a, ref))
            if os.path.exists(b):
                yield a, b

    def report(ok, what, a):
        nonlocal bad, checked
        checked += 1
        bad += 0 if ok else 1
        print("%s  %s  %s" % ("same   " if ok else "DIFFERS", what, os.path.relpath(a, ref)))

    for a, b in pairs("kmers/*.kmers.bin"):
        ra, rb = sorted_records(a), sorted_records(b)
        report(ra.shape == rb.shape and bool((ra == rb).all()), "records (multiset)", a)
    for pattern in ("stats/*.stat.txt", "vectors/*.vec", "vectors/*.breadth", "sequences/*.seq.fasta"):
        for a, b in pairs(pattern):
            if pattern.endswith(".seq.fasta"):                       # sequences come in thread order: compare the sets
                def seqs(p):
                    out, cur = [], []
                    for ln in open(p):
                        if ln.startswith(">"):
                            if cur:
                                out.append("".join(cur))
                            cur = []
                        else:
                            cur.append(ln.strip())
                    if cur:
                        out.append("".join(cur))
                    return sorted(out)
                report(seqs(a) == seqs(b), "sequences (set)", a)
            else:
                report(open(a, "rb").read() == open(b, "rb").read(), "bytes", a)
    for a, b in pairs("components.bin"):
        report(components(a) == components(b), "components (set)", a)
    ma = sorted(glob.glob(os.path.join(ref, "matrices", "*_original_order.txt")))
    mb = sorted(glob.glob(os.path.join(new, "matrices", "*_original_order.txt")))
    if ma and mb:
        report(matrix(ma[-1]) == matrix(mb[-1]), "distance matrix (cells)", ma[-1])
    print("%d compared, %d differ" % (checked, bad))
    return 1 if bad or not checked else 0


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    sys.exit(main(sys.argv[1], sys.argv[2]))
