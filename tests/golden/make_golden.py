"""Regenerates tests/golden/config1.json (BASELINE config 1 known answers).

The reference cannot be run (no JVM, lib/itmo-assembler.jar missing), so the vectors come from
the two independent restatements of the cited reference lines, which must agree:
oracle/oracle.py (numpy) and oracle/ref_cpu.c (threads + striped maps, like the reference).
Inputs: tests/golden/inputs/meta_test_{1,2,3}.fa = /root/reference/test_data/meta_test_*.fa
(verbatim copies of the reference's own fixtures).

tests/golden/meta_test_matrix.txt is NOT generated: it is a verbatim copy of the reference's checked-in result
/root/reference/test_data/meta_test_matrix.txt (the matrix-builder output for the same three files), the fixture that
pins the oracle (tests/test_oracle.py::test_reference_matrix_golden).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc          # noqa: E402
from tests import _oracle_c               # noqa: E402

out = {}
for n in (1, 2, 3):
    path = os.path.join(ROOT, "tests", "golden", "inputs", "meta_test_%d.fa" % n)
    reads = orc.parse_reads(path)
    counts = orc.count_reads(reads, 31)
    rec = orc.kmers_bin(counts, 1, 31)
    bases, offsets = _oracle_c.parse_file(path)
    c_rec, c_hist, c_distinct, _ = _oracle_c.count(bases, offsets, 31, 1, P=8)
    assert c_rec == rec and c_distinct == len(counts), "the two oracles disagree"
    out["meta_test_%d" % n] = {
        "reads": len(reads),
        "kmer_instances": sum(len(r) - 30 for r in reads),
        "distinct": len(counts),
        "kept_b1": len(rec) // 10,
        "kmers_bin_bytes": len(rec),
        "sha256_sorted_records": orc.sha256_hex(rec),
        "max_count": max(counts.values()),
        "hist": {str(c): v for c, v in orc.histogram(counts).items()},
        "stat_txt_sha256": orc.sha256_hex(orc.stat_txt(counts).encode()),
    }
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "config1.json"), "w"), indent=1, sort_keys=True)
print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "hist"} for k, v in out.items()}, indent=1))
