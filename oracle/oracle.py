"""CPU oracle (numpy / pure Python) for MetaFast's k-mer counting hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``metafast_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker.

It restates, line by line, the semantics of the reference Java code.  Citation
prefixes: ``src/`` = /root/reference/src/, ``[itmo]/`` = the ITMO assembler
source jar (/root/reference/lib/itmo-assembler-src.jar!/ru/ifmo/genetics/).

Parity status: PINNED, transitively, by the reference's own fixture.  The reference holds no unit test for this path
 and no JVM exists in the build image, but it ships one golden output:
test_data/meta_test_matrix.txt, the Bray-Curtis matrix of `matrix-builder -k 31` over test_data/meta_test_{1,2,3}.fa
(README.md:90-99).  This module restates the whole pipeline behind it (kmer-counter-many -> seq-builder ->
component-cutter -> features-calculator -> dist-matrix-calculator, `matrix_builder` below) and reproduces the three
distances to the last bit of the doubles (tests/test_oracle.py::test_reference_matrix_golden; the same test shows that
an off-by-one in the -b filter moves all of them).  What the fixture does NOT pin: FASTQ parsing (the fixture is
FASTA), saturation at 32767 and minSeqLen effects (not reached by these inputs), output ORDER (hash-map order in the
reference), the set tools.  Those stay pinned by (a) hand-derived micro known-answers and (b) agreement with the
independent C restatement ``oracle/ref_cpu.c``.  For k > 31 (128-bit keys) the reference has no behaviour at all
(src/tools/KmersCounterMain.java:70-73): parity unpinned.
"""
from __future__ import annotations

import gzip
import hashlib
import io
import os
import struct
from collections import Counter
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

MAX_COUNT = 32767  # Short.MAX_VALUE, [itmo]/utils/NumUtils.java:21-26
RECORD_SIZE = 10   # src/io/KmersLoadWorker.java:9

# [itmo]/dna/DnaTools.java:31,46-60 : A=0 G=1 C=2 T=3 (upper and lower case)
_CODE = {"A": 0, "a": 0, "G": 1, "g": 1, "C": 2, "c": 2, "T": 3, "t": 3}
_CODE_LUT = np.full(256, 255, dtype=np.uint8)
for _ch, _v in _CODE.items():
    _CODE_LUT[ord(_ch)] = _v
NUCLEOTIDES = "AGCT"  # [itmo]/dna/DnaTools.java:31


# --------------------------------------------------------------------------
# k-mer arithmetic
# --------------------------------------------------------------------------
def code(ch: str) -> int:
    """[itmo]/dna/DnaTools.java:46-64 (IUPAC codes resolve through an unseeded
    Random in the reference, so they are rejected here)."""
    try:
        return _CODE[ch]
    except KeyError:
        raise ValueError('Incorrect nucleotide char: "%s"' % ch)


def reverse_complement(kmer: int, k: int) -> int:
    """[itmo]/utils/KmerUtils.java:12-22, restated on Python ints."""
    rc = 0
    for _ in range(k):
        rc = (rc << 2) | (3 - (kmer & 3))
        kmer >>= 2
    return rc


def canonical_kmers(read: str, k: int) -> List[int]:
    """Rolling canonical k-mers of one read, in read order.

    [itmo]/dna/kmers/ShortKmer.java:122-149 (iterator), :25-31 (first k-mer),
    :68-71 (shiftRight), :54-56 (toLong = min(fw, rc)).  Works for any k
    (Python ints), which is how the k>31 extension is defined.
    """
    n = len(read)
    if n < k:
        return []
    mask = (1 << (2 * k)) - 1
    fw = 0
    rc = 0
    out = []
    for i, ch in enumerate(read):
        c = code(ch)
        fw = ((fw << 2) | c) & mask
        rc = (rc >> 2) | ((3 - c) << (2 * k - 2))
        if i >= k - 1:
            out.append(fw if fw < rc else rc)
    return out


def kmer_to_string(kmer: int, k: int) -> str:
    """[itmo]/utils/KmerUtils.java:50-57."""
    return "".join(NUCLEOTIDES[(kmer >> (2 * i)) & 3] for i in range(k - 1, -1, -1))


def _codes_of(read: str) -> np.ndarray:
    a = _CODE_LUT[np.frombuffer(read.encode("latin-1"), dtype=np.uint8)]
    if (a == 255).any():
        bad = read[int(np.argmax(a == 255))]
        raise ValueError('Incorrect nucleotide char: "%s"' % bad)
    return a


def canonical_kmers_np(reads: Sequence[str], k: int, min_len: int = 0) -> np.ndarray:
    """Vectorised version of ``canonical_kmers`` for k <= 31 (uint64 keys).

    Reads shorter than ``min_len`` are skipped (src/io/IOUtils.java:761), reads
    shorter than k yield nothing ([itmo]/dna/kmers/ShortKmer.java:123,127).
    """
    assert 1 <= k <= 31
    by_len: Dict[int, List[np.ndarray]] = {}
    for r in reads:
        if len(r) < k or len(r) < min_len:
            continue
        by_len.setdefault(len(r), []).append(_codes_of(r))
    outs = []
    for L, lst in by_len.items():
        c = np.stack(lst).astype(np.uint64)             # (n, L)
        nk = L - k + 1
        fw = np.zeros((c.shape[0], nk), dtype=np.uint64)
        rc = np.zeros((c.shape[0], nk), dtype=np.uint64)
        for j in range(k):
            w = c[:, j:j + nk]
            fw |= w << np.uint64(2 * (k - 1 - j))
            rc |= (np.uint64(3) - w) << np.uint64(2 * j)
        outs.append(np.minimum(fw, rc).ravel())
    if not outs:
        return np.zeros(0, dtype=np.uint64)
    return np.concatenate(outs)


# --------------------------------------------------------------------------
# parsers  (a3 in SURVEY.md section 8)
# --------------------------------------------------------------------------
def detect_file_format(path: str) -> str:
    """[itmo]/io/ReadersUtils.java:27-54."""
    name = os.path.basename(path).lower()
    suffix = ""
    if name.endswith(".gz"):
        suffix = ".gz"
        name = name[:-3]
    if name.endswith(".bz2"):
        suffix = ".bz2"
        name = name[:-4]
    if name.endswith(".binq"):
        return "binq" + suffix
    if name.endswith(".fastq") or name.endswith(".fq"):
        return "fastq" + suffix
    if name.endswith((".fasta", ".fa", ".fn", ".fna")):
        return "fasta" + suffix
    raise IOError("Can't detect file format for file '%s'" % name)


def _remove_extension(s: str, *exts: str) -> str:
    """[itmo]/utils/FileUtils.java:199-210 (first matching extension only)."""
    for e in exts:
        if s.lower().endswith(e.lower()):
            return s[: len(s) - len(e)]
    return s


def library_name(path: str) -> str:
    """``NamedSource.name()``: FastaReader.java:22, FastaGZReader.java:17,
    FastqReader.java:25, FastqGZReader.java:21."""
    fmt = detect_file_format(path)
    base = os.path.basename(path)
    if fmt == "fasta":
        return _remove_extension(base, ".fasta", ".fa", ".fn", ".fna")
    if fmt == "fasta.gz":
        return _remove_extension(base, ".fasta.gz", ".fa.gz", ".fn.gz", ".fna.gz")
    if fmt == "fastq":
        return _remove_extension(base, ".fastq", ".fq")
    if fmt == "fastq.gz":
        return _remove_extension(base, ".fastq.gz", ".fq.gz")
    raise IOError("format %s is out of scope" % fmt)


def _read_lines(path: str) -> List[str]:
    """BufferedReader.readLine semantics: split on \\n, \\r\\n or \\r; a final
    line without terminator is still a line."""
    opener = gzip.open if path.lower().endswith(".gz") else open
    with opener(path, "rb") as f:
        data = f.read()
    text = data.decode("latin-1")
    lines = text.replace("\r\n", "\n").replace("\r", "\n").split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    return lines


def parse_fasta(path: str) -> List[str]:
    """[itmo]/io/readers/FastaReader.java:54-108: '>' / ';' lines delimit
    records, other lines are concatenated, records containing N/n are
    dropped."""
    out: List[str] = []
    sb: List[str] = []

    def flush():
        if sb:
            s = "".join(sb)
            sb.clear()
            if "N" not in s and "n" not in s:
                out.append(s)

    for line in _read_lines(path):
        if line.startswith(">") or line.startswith(";"):
            flush()
        else:
            sb.append(line)
    flush()
    for s in out:
        _codes_of(s)  # IllegalArgumentException on anything but ACGTacgt
    return out


class IllegalQualityValue(ValueError):
    pass


def _phred(fmt: str, q: str) -> int:
    """[itmo]/io/formats/Illumina.java:7-12, Sanger.java:7-12."""
    c = ord(q)
    lo = 64 if fmt == "illumina" else 33
    if c < lo or c > 126:
        raise IllegalQualityValue('Invalid quality code char: "%s"' % q)
    return c - lo


def _fastq_records(lines: List[str]) -> Iterator[Tuple[str, str]]:
    """[itmo]/io/readers/FastqReader.java:53-66,84-110."""
    i = 0
    n = len(lines)

    def next_data_line() -> Optional[str]:
        nonlocal i
        while i < n and len(lines[i]) == 0:
            i += 1
        if i >= n:
            return None
        s = lines[i]
        i += 1
        if not (s.startswith("@") or s.startswith("+")):
            raise RuntimeError("Unknown structure of fastq file!")
        if i >= n:
            raise RuntimeError("Unexpected end of file.")
        s = lines[i]
        i += 1
        return s

    while True:
        data = next_data_line()
        if data is None:
            return
        qual = next_data_line()
        if qual is None:
            raise RuntimeError("Unexpected end of file.")
        if len(data) != len(qual):
            raise RuntimeError("Bad DnaQ record: length of chars and quality is not the same.")
        yield data, qual


def _fastq_keep(fmt: str, data: str, qual: str) -> bool:
    """FastqReader.java:70-79 + DnaQBuilder.java:32-35 + DnaQ.java:140-150 +
    FastaReaderFromXQSource.java:62-69: kept iff every position has
    (phred & 63) != 0; N/n/. store phred 0."""
    good = True
    for ch, q in zip(data, qual):       # every position is parsed (errors throw)
        if ch in "Nn.":
            good = False
            continue
        code(ch)
        if (_phred(fmt, q) & 63) == 0:
            good = False
    return good


def determine_quality_format(lines: List[str], head: int = 1000) -> str:
    """[itmo]/io/ReadersUtils.java:57-77: Illumina unless parsing the first
    1000 records as Illumina throws IllegalQualityValueException."""
    try:
        for idx, (data, qual) in enumerate(_fastq_records(lines)):
            if idx >= head:
                break
            for ch, q in zip(data, qual):
                if ch in "Nn.":
                    continue
                code(ch)
                _phred("illumina", q)
    except IllegalQualityValue:
        return "sanger"
    return "illumina"


def parse_fastq(path: str) -> List[str]:
    lines = _read_lines(path)
    fmt = determine_quality_format(lines)
    return [d for d, q in _fastq_records(lines) if _fastq_keep(fmt, d, q)]


def parse_reads(path: str) -> List[str]:
    """[itmo]/io/ReadersUtils.java:85-102."""
    fmt = detect_file_format(path)
    if fmt in ("fasta", "fasta.gz"):
        return parse_fasta(path)
    if fmt in ("fastq", "fastq.gz"):
        return parse_fastq(path)
    raise IOError("format %s is out of scope" % fmt)


# --------------------------------------------------------------------------
# sample discovery (a1, a2)
# --------------------------------------------------------------------------
def group_samples(paths: Sequence[str]) -> List[Tuple[str, List[str]]]:
    """src/tools/KmersCounterForManyFilesMain.java:73-108 and
    src/tools/KmersCounterMain.java:122-137.  Returns [(sample name, files)]."""
    files = sorted(paths)
    names = [library_name(f) for f in files]
    out = []
    i = 0
    while i < len(files):
        if i + 1 < len(files) and (
            (names[i].endswith("_r1") and names[i + 1].endswith("_r2"))
            or (names[i].endswith("_R1") and names[i + 1].endswith("_R2"))
        ):
            out.append((names[i][:-3], [files[i], files[i + 1]]))
            i += 2
        else:
            out.append((names[i], [files[i]]))
            i += 1
    return out


# --------------------------------------------------------------------------
# count + emit (a6, a7, a8)
# --------------------------------------------------------------------------
def count_reads(reads: Iterable[str], k: int, min_len: int = 0) -> Dict[int, int]:
    """src/io/IOUtils.java:756-769 + Long2ShortHashMap.addAndBound: exact
    multiset count saturated at 32767."""
    reads = list(reads)
    if k <= 31:
        keys = canonical_kmers_np(reads, k, min_len)
        u, c = np.unique(keys, return_counts=True)
        return {int(a): int(min(b, MAX_COUNT)) for a, b in zip(u, c)}
    cnt: Counter = Counter()
    for r in reads:
        if len(r) >= min_len:
            cnt.update(canonical_kmers(r, k))
    return {a: min(b, MAX_COUNT) for a, b in cnt.items()}


def read_stats(reads: Sequence[str], min_len: int = 0) -> Tuple[int, int, int, int]:
    """totalSeq, goodSeq, totalLen, goodLen of src/io/IOUtils.java:756-769."""
    tot = len(reads)
    tot_len = sum(len(r) for r in reads)
    good = [r for r in reads if len(r) >= min_len]
    return tot, len(good), tot_len, sum(len(r) for r in good)


def histogram(counts: Dict[int, int]) -> Dict[int, int]:
    """QuickQuantitativeStatistics over ALL entries, src/io/IOUtils.java:59."""
    return dict(sorted(Counter(counts.values()).items()))


def key_bytes(k: int) -> int:
    return 8 if k <= 31 else 16


def kmers_bin(counts: Dict[int, int], threshold: int, k: int = 31) -> bytes:
    """src/io/IOUtils.java:61-65: BE int64 key + BE int16 count for
    count > threshold.  Ascending key order (the reference's order is hash-table
    iteration order and not part of the contract).  For
    k > 31 the key is 16 bytes big-endian (this repo's extension)."""
    kb = key_bytes(k)
    out = bytearray()
    for key in sorted(counts):
        c = counts[key]
        if c > threshold:
            out += key.to_bytes(kb, "big") + struct.pack(">h", c)
    return bytes(out)


def stat_txt(counts: Dict[int, int]) -> str:
    """[itmo]/statistics/QuickQuantitativeStatistics.java:38-55,65-72 with the
    header of src/io/IOUtils.java:69.  println(header); println(toString())."""
    s = "# k-mer frequency\tnumber of such k-mers\n"
    for c, n in histogram(counts).items():
        s += "%d\t%d\n" % (c, n)
    return s + "\n"


def load_kmers_bin(data: bytes, k: int = 31) -> List[Tuple[int, int]]:
    """src/io/KmersLoadWorker.java:16-34: 10-byte BE records (signed short)."""
    rs = key_bytes(k) + 2
    if len(data) % rs:
        raise RuntimeError("BAD division by work range")
    out = []
    for i in range(0, len(data), rs):
        key = int.from_bytes(data[i:i + rs - 2], "big", signed=False)
        (freq,) = struct.unpack(">h", data[i + rs - 2:i + rs])
        out.append((key, freq))
    return out


def sorted_records(data: bytes, k: int = 31) -> bytes:
    rs = key_bytes(k) + 2
    recs = [data[i:i + rs] for i in range(0, len(data), rs)]
    recs.sort()
    return b"".join(recs)


def sha256_hex(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def load_kmers(datas: Sequence[bytes], threshold: int, k: int = 31) -> Dict[int, int]:
    """src/io/IOUtils.java:237-258,369-401: map[key] = sat_add16(map[key], freq)
    for records with freq > threshold."""
    hm: Dict[int, int] = {}
    for d in datas:
        for key, freq in load_kmers_bin(d, k):
            if freq > threshold:
                hm[key] = min(hm.get(key, 0) + freq, MAX_COUNT)
    return hm


# --------------------------------------------------------------------------
# set algebra over .kmers.bin files
# --------------------------------------------------------------------------
def _java_short(x: int) -> int:
    """(short) cast: wraps to 16 bits, signed."""
    x &= 0xFFFF
    return x - 0x10000 if x >= 0x8000 else x


def _get(hm: Dict[int, int], key: int) -> int:
    """Long2ShortHashMap.get ([itmo]/structures/map/Long2ShortHashMap.java:160-175): -1 when absent."""
    return hm.get(key, -1)


def _get_with_zero(hm: Dict[int, int], key: int) -> int:
    """Long2ShortHashMap.getWithZero (:178-183): an absent key AND a stored -1 read as 0."""
    v = _get(hm, key)
    return 0 if v == -1 else v


def _records_of(entries: Iterable[Tuple[int, int]]) -> bytes:
    """10-byte BE records in ascending key order (the reference writes iteration order; compare sorted)."""
    return b"".join(struct.pack(">Qh", key, v) for key, v in sorted(entries))


def filter_and_print_kmers(hm: Dict[int, int], filter_hm: Dict[int, int], threshold: int, filter_threshold: int) -> bytes:
    """src/io/IOUtils.java:101-123."""
    return _records_of((key, v) for key, v in hm.items() if v > threshold and _get_with_zero(filter_hm, key) > filter_threshold)


def kmers_filter(inputs: Sequence[bytes], filters: Sequence[bytes], b: int = 1, max_thresh: int = 0) -> List[Tuple[int, bytes]]:
    """src/tools/KmersFilter.java:94-110: per input file -> (hm.size(), filtered records)."""
    filter_hm = load_kmers(filters, b)
    out = []
    for data in inputs:
        hm = load_kmers([data], b)
        out.append((len(hm), filter_and_print_kmers(hm, filter_hm, b, max_thresh * len(filters))))
    return out


def unique_kmers_multi(inputs: Sequence[bytes], filters: Sequence[bytes], b: int = 1, min_samples: int = 1,
                       max_samples: int = 1) -> Tuple[int, Dict[int, bytes]]:
    """src/tools/UniqueKmersMultipleSamplesFinder.java:97-148 -> (hm.size(), {i: records of filtered_i.kmers.bin})."""
    hm: Dict[int, int] = {}
    hm_cnt: Dict[int, int] = {}
    for data in inputs:
        for key, value in load_kmers([data], b).items():
            if value > b:
                hm[key] = _java_short(_get_with_zero(hm, key) + value)              # :107-108, (short) wraps
                hm_cnt[key] = _java_short(_get_with_zero(hm_cnt, key) + 1)          # :109
    for data in filters:
        for key, value in load_kmers([data], b).items():
            if value > b and _get(hm, key) > b:                                      # :127-129
                hm[key] = 0
    return len(hm), {i: filter_and_print_kmers(hm, hm_cnt, b, i - 1) for i in range(min_samples, max_samples + 1)}


def kmers_samples_counter(inputs: Sequence[bytes], b: int = 1) -> Tuple[int, bytes, str]:
    """src/tools/KmersSamplesCounter.java:90-119 -> (hm.size(), n_samples.kmers.bin, n_samples.stat.txt)."""
    hm = {key: 0 for key in load_kmers(inputs, b)}                                   # loadKmers + resetValues
    for data in inputs:
        for key, value in load_kmers([data], b).items():
            if value > b:
                hm[key] = _java_short(_get_with_zero(hm, key) + 1)
    records = _records_of((key, v) for key, v in hm.items() if v > 0)                # printKmers(hm, 0, ...)
    return len(hm), records, stat_txt(hm)


# --------------------------------------------------------------------------
# seq-builder: simple paths of the de Bruijn graph of the counted k-mers
# --------------------------------------------------------------------------
def _shift_right(fw: int, nuc: int, k: int) -> int:
    """ShortKmer.shiftRight ([itmo]/dna/kmers/ShortKmer.java:68-71), forward strand only"""
    return ((fw << 2) | nuc) & ((1 << (2 * k)) - 1)


def _shift_left(fw: int, nuc: int, k: int) -> int:
    """ShortKmer.shiftLeft (:89-92)"""
    return (fw >> 2) | (nuc << (2 * k - 2))


def _canon(fw: int, k: int) -> int:
    return min(fw, reverse_complement(fw, k))


def _left_nucleotide(hm: Dict[int, int], fw: int, k: int, thr: int) -> int:
    """HashMapOperations.getLeftNucleotide (src/algo/HashMapOperations.java:13-29): the unique nucleotide that extends
    the k-mer to the left into a k-mer with count > thr; -1 = none, -2 = several"""
    ans = -1
    for nuc in range(4):
        if _get(hm, _canon(_shift_left(fw, nuc, k), k)) > thr:
            if ans > -1:
                return -2
            ans = nuc
    return ans


def _right_nucleotide(hm: Dict[int, int], fw: int, k: int, thr: int) -> int:
    """HashMapOperations.getRightNucleotide (:31-47)"""
    ans = -1
    for nuc in range(4):
        if _get(hm, _canon(_shift_right(fw, nuc, k), k)) > thr:
            if ans > -1:
                return -2
            ans = nuc
    return ans


def seq_builder(hm: Dict[int, int], k: int, freq_threshold: int, len_threshold: int) -> List[Tuple[str, int, int, int]]:
    """SequencesFinders.thresholdStrategy + AddSequencesShiftingRightTask (src/algo/SequencesFinders.java:13-31,
    src/algo/AddSequencesShiftingRightTask.java:39-123) -> [(sequence, av_weight, min_weight, max_weight)], in ascending
    order of (start k-mer key, orientation) -- the reference fills a concurrent deque in thread order, so only the SET
    is defined; both orientations of one key are visited by one thread, forward first (:52-53), which makes the
    `used` rule for start == end deterministic."""
    out = []
    used = set()
    for key in sorted(hm):
        if hm[key] <= freq_threshold:
            continue
        for fw in (key, reverse_complement(key, k)):
            nuc = _left_nucleotide(hm, fw, k, freq_threshold)
            is_left = nuc < 0
            if not is_left:
                if _right_nucleotide(hm, _shift_left(fw, nuc, k), k, freq_threshold) < 0:
                    is_left = True
            if not is_left:
                continue
            value = _get_with_zero(hm, _canon(fw, k))                               # processSequence :75-122
            seq = [kmer_to_string(fw, k)]
            weight, lo, hi = value, value, value
            cur = fw
            while True:
                r = _right_nucleotide(hm, cur, k, freq_threshold)
                if r < 0:
                    break
                nxt = _shift_right(cur, r, k)
                if _left_nucleotide(hm, nxt, k, freq_threshold) < 0:
                    break
                cur = nxt
                seq.append("AGCT"[r])
                value = _get_with_zero(hm, _canon(cur, k))
                weight += value
                lo, hi = min(lo, value), max(hi, value)
            text = "".join(seq)
            if len(text) < len_threshold:
                continue
            st, end = _canon(fw, k), _canon(cur, k)
            if st > end:
                continue
            if st == end:
                if st in used:
                    continue
                used.add(st)
            out.append((text, weight // (len(text) - k + 1), lo, hi))
    return out


def sequences_fasta(seqs: Sequence[Tuple[str, int, int, int]]) -> str:
    """Sequence.printSequences (src/structures/Sequence.java:26-37) + FastaDedicatedWriter.writeData
    ([itmo]/io/writers/FastaDedicatedWriter.java:34-52): 70 bases per line"""
    out = []
    for i, (text, av, lo, hi) in enumerate(seqs, 1):
        out.append(">%d length=%d av_weight=%d min_weight=%d max_weight=%d\n" % (i, len(text), av, lo, hi))
        for j in range(0, len(text), 70):
            out.append(text[j:j + 70] + "\n")
    return "".join(out)


def seq_builder_distribution(hm: Dict[int, int], stat_len: int = 1024) -> str:
    """SeqBuilderMain.runImpl :83-101 + dumpStat :170-176: "<count> <number of k-mers>" for count 1..1023 (last bin = rest)"""
    stat = [0] * stat_len
    for v in hm.values():
        stat[min(v, stat_len - 1)] += 1
    return "".join("%d %d\n" % (i, stat[i]) for i in range(1, stat_len))


# --------------------------------------------------------------------------
# features-calculator (a10-a13)
# --------------------------------------------------------------------------
def load_components(data: bytes, k: int = 31) -> List[Tuple[int, List[int]]]:
    """src/structures/ConnectedComponent.java:95-122: BE int32 n; n x
    {BE int32 size; BE int64 weight; size x BE int64 key}."""
    kb = key_bytes(k)
    (n,) = struct.unpack_from(">i", data, 0)
    off = 4
    out = []
    for _ in range(n):
        size, weight = struct.unpack_from(">iq", data, off)
        off += 12
        keys = [int.from_bytes(data[off + kb * j: off + kb * (j + 1)], "big") for j in range(size)]
        off += kb * size
        out.append((weight, keys))
    return out


def save_components(comps: Sequence[Tuple[int, Sequence[int]]], k: int = 31) -> bytes:
    """src/structures/ConnectedComponent.java:80-93."""
    kb = key_bytes(k)
    out = bytearray(struct.pack(">i", len(comps)))
    for weight, keys in comps:
        out += struct.pack(">iq", len(keys), weight)
        for key in keys:
            out += int(key).to_bytes(kb, "big")
    return bytes(out)


_I64_MAX = (1 << 63) - 1


def _java_add_and_bound64(value: int, inc: int) -> int:
    """[itmo]/utils/NumUtils.java:27-32 with Java's wrapping long arithmetic."""
    lim = _I64_MAX - inc
    lim = (lim + (1 << 63)) % (1 << 64) - (1 << 63)   # wrap to int64
    if value > lim:
        return _I64_MAX
    return value + inc


def presence_for_kmers(component_keys: Iterable[int], records: Iterable[Tuple[int, int]]) -> Dict[int, int]:
    """src/tools/FeaturesCalculatorMain.java:97-103 (seed with 0) and
    src/io/IOUtils.java:577-588 (if contains: addAndBound(kmer, freq))."""
    acc = {key: 0 for key in component_keys}
    for key, freq in records:
        if key in acc:
            acc[key] = _java_add_and_bound64(acc[key], freq)
    return acc


def presence_for_reads(component_keys: Iterable[int], reads: Iterable[str], k: int) -> Dict[int, int]:
    """src/io/IOUtils.java:806-825 (no minSeqLen on this path)."""
    acc = {key: 0 for key in component_keys}
    for r in reads:
        for key in canonical_kmers(r, k):
            if key in acc:
                acc[key] = _java_add_and_bound64(acc[key], 1)
    return acc


def features(comps: Sequence[Tuple[int, Sequence[int]]], acc: Dict[int, int], threshold: int = 0,
             selected: Optional[Dict[int, int]] = None) -> Tuple[List[int], List[float], List[int], List[int]]:
    """src/tools/FeaturesCalculatorMain.java:186-204.  Returns
    (vec, breadth, found, cnt)."""
    vec, breadth, founds, cnts = [], [], [], []
    for _w, keys in comps:
        s = 0
        found = 0
        cnt = 0
        for key in keys:
            if selected is None or selected.get(key, 0) > 0:
                v = acc.get(key, 0)
                if v > threshold:
                    s += v
                    s = (s + (1 << 63)) % (1 << 64) - (1 << 63)  # Java long wraps
                    found += 1
                cnt += 1
        vec.append(s)
        breadth.append(float("nan") if cnt == 0 else found / cnt)
        founds.append(found)
        cnts.append(cnt)
    return vec, breadth, founds, cnts


def java_double_to_string(d: float) -> str:
    """java.lang.Double.toString as of JDK 19+ (shortest repr that round-trips;
    older JDKs occasionally print one digit more).  Used for ``.breadth``
    (src/tools/FeaturesCalculatorMain.java:225-230)."""
    if d != d:
        return "NaN"
    if d in (float("inf"), float("-inf")):
        return "Infinity" if d > 0 else "-Infinity"
    if d == 0:
        return "-0.0" if str(d).startswith("-") else "0.0"
    sign = "-" if d < 0 else ""
    r = repr(abs(d))
    # digits and decimal exponent from Python's shortest repr
    if "e" in r:
        mant, e = r.split("e")
        exp10 = int(e)
    else:
        mant, exp10 = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    digits = (ip + fp).lstrip("0")
    # position of decimal point relative to the first significant digit
    lead_zeros = len(ip + fp) - len((ip + fp).lstrip("0"))
    point = len(ip) - lead_zeros + exp10          # value = 0.DIGITS * 10^point
    digits = digits.rstrip("0") or "0"
    a = abs(d)
    if 1e-3 <= a < 1e7:
        if point <= 0:
            s = "0." + "0" * (-point) + digits
        elif point >= len(digits):
            s = digits + "0" * (point - len(digits)) + ".0"
        else:
            s = digits[:point] + "." + digits[point:]
        return sign + s
    frac = digits[1:] or "0"
    return "%s%s.%sE%d" % (sign, digits[0], frac, point - 1)


def vec_text(vec: Sequence[int]) -> str:
    return "".join("%d\n" % v for v in vec)


def breadth_text(breadth: Sequence[float]) -> str:
    return "".join(java_double_to_string(b) + "\n" for b in breadth)


# --------------------------------------------------------------------------
# component-cutter and dist-matrix-calculator: the rest of the reference's default `matrix-builder` pipeline.
# NOT on the accelerated path; restated here only because the reference's single golden output,
# test_data/meta_test_matrix.txt, is the result of the WHOLE pipeline -- with these two stages the oracle can be
# pinned against it (tests/test_oracle.py::test_reference_matrix_golden).
# --------------------------------------------------------------------------
def possible_neighbours(key: int, k: int) -> List[int]:
    """KmerOperations.possibleNeighbours (src/algo/KmerOperations.java:9-26): the canonical forms of the 4 right and
    4 left extensions of a (canonical) k-mer"""
    out = []
    for nuc in range(4):
        out.append(_canon(_shift_right(key, nuc, k), k))
        out.append(_canon(_shift_left(key, nuc, k), k))
    return out


def _find_all_components(hm: Dict[int, int], k: int) -> List[List[int]]:
    """ComponentsBuilder.findAllComponents + bfs (src/algo/ComponentsBuilder.java:198-269): connected components of
    the k-mers of `hm` under the neighbour relation.  The reference walks `hm` in hash-map order and marks visited
    k-mers by negating their value; a component larger than b2 is still walked to its end (:244-262), so the SET of
    components does not depend on the order"""
    seen = set()
    comps = []
    for start in sorted(hm):
        if start in seen:
            continue
        seen.add(start)
        comp = [start]
        head = 0
        while head < len(comp):
            kmer = comp[head]
            head += 1
            for nb in possible_neighbours(kmer, k):
                if nb in hm and nb not in seen:
                    seen.add(nb)
                    comp.append(nb)
        comps.append(comp)
    return comps


def component_cutter(hm: Dict[int, int], k: int, b1: int = 1000, b2: int = 10000) -> List[Tuple[int, List[int], int]]:
    """ComponentsBuilder.splitStrategy / run / Task.run (src/algo/ComponentsBuilder.java:24-31,58-84,157-181) ->
    [(weight, keys, usedFreqThreshold)]: components of fewer than b1 k-mers are dropped, b1..b2 are kept, larger ones
    are split again among their k-mers of value >= threshold + 1 (`nextHM`, :246-260).  Sorted as
    ConnectedComponent.compareTo (src/structures/ConnectedComponent.java:125-136: threshold ascending, weight
    descending, size descending); ties and the order of the keys inside a component are hash-map / thread order in
    the reference, here ascending smallest key / ascending key"""
    ans = []
    work = [(hm, 1)]
    while work:
        cur, thr = work.pop()
        for comp in _find_all_components(cur, k):
            if len(comp) < b1:
                continue
            if len(comp) <= b2:
                ans.append((sum(cur[x] for x in comp), sorted(comp), thr))
            else:
                work.append(({x: cur[x] for x in comp if cur[x] >= thr + 1}, thr + 1))
    ans.sort(key=lambda c: (c[2], -c[0], -len(c[1]), c[1][0]))
    return ans


def components_stat_txt(comps: Sequence[Tuple[int, Sequence[int], int]]) -> str:
    """src/algo/ComponentsBuilder.java:144-152"""
    out = ["# component.no\tcomponent.size\tcomponent.weight\tusedFreqThreshold\n"]
    for i, (weight, keys, thr) in enumerate(comps, 1):
        out.append("%d\t%d\t%d\t%d\n" % (i, len(keys), weight, thr))
    return "".join(out)


def bray_curtis(v1: Sequence[float], v2: Sequence[float]) -> float:
    """DistanceMatrixCalculatorMain.brayCurtisDistance (src/tools/DistanceMatrixCalculatorMain.java:140-153), in
    doubles and in the reference's summation order"""
    sumdiff = 0.0
    total = 0.0
    for a, b in zip(v1, v2):
        a, b = float(a), float(b)
        sumdiff += abs(a - b)
        total += abs(a) + abs(b)
    return sumdiff / total


def matrix_builder(paths: Sequence[str], k: int = 31, b: int = 1, min_seq_len: int = 100, b1: int = 1000,
                   b2: int = 10000) -> Tuple[List[str], List[List[float]], dict]:
    """``matrix-builder -k K -b B -l L -i paths`` up to the distance matrix in original order
    (src/tools/DistanceMatrixBuilderMain.java:88-137,170-175): kmer-counter-many -> seq-builder-many (one run per
    .kmers.bin, src/tools/SeqBuilderForManyFilesMain.java:80-92) -> component-cutter (loadReads over ALL sequence
    files with minSeqLen, src/tools/ComponentCutterMain.java:81-82) -> features-calculator on the .kmers.bin files
    (threshold 0) -> Bray-Curtis.  Returns (sample names, matrix, intermediates)."""
    counted = kmer_counter_many(paths, k, b)
    names = list(counted)
    seqs = {}
    for name in names:
        hm = load_kmers([counted[name][0]], b, k)                          # SeqBuilderMain.java:79-80
        seqs[name] = seq_builder(hm, k, b, min_seq_len)
    seq_hm = count_reads([s[0] for name in names for s in seqs[name]], k, min_seq_len)
    comps3 = component_cutter(seq_hm, k, b1, b2)
    comps = [(w, keys) for w, keys, _thr in comps3]
    all_keys = [key for _w, keys in comps for key in keys]
    vecs = {}
    for name in names:
        acc = presence_for_kmers(all_keys, load_kmers_bin(counted[name][0], k))
        vecs[name] = features(comps, acc, 0)[0]
    n = len(names)
    matrix = [[0.0] * n for _ in range(n)]
    for i in range(n):
        for j in range(i + 1, n):
            matrix[i][j] = matrix[j][i] = bray_curtis(vecs[names[i]], vecs[names[j]])
    return names, matrix, {"counted": counted, "sequences": seqs, "sequence_kmers": seq_hm, "components": comps3,
                           "vectors": vecs}


def java_format_fixed(d: float, prec: int = 4) -> str:
    """java.util.Formatter "%.<prec>f" for a double: the shortest round-trip digits (FormattedFloatingDecimal), rounded
    HALF_UP -- not C's exact-binary half-even (0.125 -> "0.13" in Java, "0.12" in C)"""
    from decimal import Decimal, ROUND_HALF_UP
    q = Decimal(1).scaleb(-prec)
    return format(Decimal(repr(float(d))).quantize(q, rounding=ROUND_HALF_UP), "f")


def heatmap_order(matrix: Sequence[Sequence[float]]) -> List[int]:
    """FullHeatMap.clusterObjects + renumber (src/algo/FullHeatMap.java:218-289,321-333): average-linkage clustering,
    the first minimum in (i, j) order is merged into slot i; the new order of the samples = leaves left to right"""
    n = len(matrix)
    groups: List[Optional[List[int]]] = [[i] for i in range(n)]

    def gdist(g1, g2):
        total = 0.0
        for x in g1:
            for y in g2:
                total += matrix[x][y]
        return total / len(g1) / len(g2)

    for _ in range(n - 1):
        best, bi, bj = float("inf"), -1, -1
        for i in range(n):
            for j in range(i + 1, n):
                if groups[i] is not None and groups[j] is not None:
                    dij = gdist(groups[i], groups[j])
                    if dij < best:
                        best, bi, bj = dij, i, j
        groups[bi] = groups[bi] + groups[bj]
        groups[bj] = None
    return [g for g in groups if g is not None][0] if n else []


def matrix_txt(matrix: Sequence[Sequence[float]], names: Optional[Sequence[str]], perm: Optional[Sequence[int]] = None,
               fmt: str = "%.4f") -> str:
    """DistanceMatrixCalculatorMain.printMatrix (src/tools/DistanceMatrixCalculatorMain.java:91-121); fmt "%.<N>f" or
    "%s" (Double.toString)"""
    n = len(matrix)
    at = (lambda i: perm[i]) if perm is not None else (lambda i: i)
    cell = java_double_to_string if fmt == "%s" else (lambda d: java_format_fixed(d, int(fmt[2:-1])))
    out = []
    if names is not None:
        out.append("#" + "".join("\t" + names[at(i)] for i in range(n)) + "\n")
    for i in range(n):
        row = "\t".join(cell(matrix[at(i)][at(j)]) for j in range(n))
        out.append((names[at(i)] + "\t" if names is not None else "") + row + "\n")
    return "".join(out)


def load_matrix_txt(text: str) -> Dict[Tuple[str, str], float]:
    """Parses DistanceMatrixCalculatorMain.printMatrix output (src/tools/DistanceMatrixCalculatorMain.java:91-121)
    into {(row name, column name): value}; the heat-map step may have permuted rows and columns"""
    lines = [ln for ln in text.splitlines() if ln.strip()]
    cols = lines[0].split("\t")[1:]
    out = {}
    for ln in lines[1:]:
        cells = ln.split("\t")
        for c, v in zip(cols, cells[1:]):
            out[(cells[0], c)] = float(v)
    return out


# --------------------------------------------------------------------------
# whole-tool restatements
# --------------------------------------------------------------------------
def kmer_counter_many(paths: Sequence[str], k: int, b: int = 1) -> Dict[str, Tuple[bytes, str, Dict[int, int]]]:
    """``kmer-counter-many -k K -b B -i paths`` -> {sample: (kmers.bin bytes
    (key-sorted), stat.txt text, counts)}."""
    if k <= 0 or k > 31:
        # src/tools/KmersCounterMain.java:66-73
        raise SystemExit(1)
    out = {}
    for name, files in group_samples(paths):
        reads: List[str] = []
        for f in files:
            reads += parse_reads(f)
        counts = count_reads(reads, k, 0)
        out[name] = (kmers_bin(counts, b, k), stat_txt(counts), counts)
    return out
