"""mfkc_cli on a FASTA record longer than its batch buffer (the reference takes records of any length).
In a file of its own, last of the GPU files: written after the round's last GPU session, so it has not run on a GPU yet."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from tests.conftest import ROOT

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "metafast_b200", "bin", "mfkc_cli")


def test_cli_record_longer_than_the_batch_buffer(built, tmp_path):
    rng = np.random.default_rng(5)
    seq = lambda n: "".join(rng.choice(list("ACGT"), n))
    genome = seq(30000)
    recs = [genome[:400], genome, genome[100:700], genome[5000:20000], seq(90)]
    fa = tmp_path / "contigs.fa"
    fa.write_text("".join(">c%d\n" % i + "\n".join(r[j:j + 70] for j in range(0, len(r), 70)) + "\n" for i, r in enumerate(recs)))
    env = dict(os.environ, MFKC_CLI_BATCH_BASES="2000", MFKC_READER_CHUNK="4096")     # every record but the last outgrows the buffer
    wd = tmp_path / "wd"
    r = subprocess.run([CLI, "-t", "kmer-counter-many", "-k", "31", "-b", "1", "-i", str(fa), "-w", str(wd)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    assert r.returncode == 0, r.stderr
    counts = orc.count_reads(recs, 31)
    assert open(wd / "kmers" / "contigs.kmers.bin", "rb").read() == orc.kmers_bin(counts, 1, 31)
    assert "5 reads added" in r.stderr
