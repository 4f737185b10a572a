"""CPU tier: bgsa_b200/convert (host C, no GPU) against the reference's own convert tool
(original/BGSA_CPU/convert.c, banded/BGSA_CPU/convert.c; binaries oracle/_ref/convert_int16|int8 built
from the unmodified sources where those are available) and against hand-derived expectations."""
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
CONVERT = ROOT / "bgsa_b200" / "convert"
REF = ROOT / "oracle" / "_ref"


def _run(exe, args, cwd):
    res = subprocess.run([str(exe)] + [str(a) for a in args], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=120)
    assert res.returncode == 0, res.stdout
    return res.stdout


@pytest.fixture(scope="module", autouse=True)
def built():
    if not CONVERT.exists():
        subprocess.check_call(["make", "-s", "-C", str(ROOT), "bgsa_b200/convert"])


def _fasta(rng, n, wrap, crlf_free=True):
    lines, seqs = [], []
    for i in range(n):
        ln = int(rng.integers(1, 200))
        s = "".join(rng.choice(list("ACGTN"), ln))
        seqs.append(s)
        lines.append(f">read{i} some description | x=@{i}")
        lines += [s[j:j + wrap] for j in range(0, ln, wrap)]
    return "\n".join(lines) + "\n", seqs


def _fastq(rng, n, at_in_quality=False):
    lines, seqs = [], []
    for i in range(n):
        ln = int(rng.integers(1, 150))
        s = "".join(rng.choice(list("ACGTN"), ln))
        q = "".join(rng.choice(list("IIHG#5+<" + ("@" if at_in_quality else "")), ln))
        seqs.append(s)
        lines += [f"@read{i}/1", s, "+", q]
    return "\n".join(lines) + "\n", seqs


def test_fasta_expected_and_reference(tmp_path):
    rng = np.random.default_rng(1)
    text, seqs = _fasta(rng, 50, 60)
    (tmp_path / "in.fa").write_text(text)
    _run(CONVERT, ["-f", "in.fa", "-o", "ours.txt"], tmp_path)
    assert (tmp_path / "ours.txt").read_text() == "\n".join(seqs) + "\n"
    if (REF / "convert_int16").exists():
        _run(REF / "convert_int16", ["-f", "in.fa", "-o", "ref.txt"], tmp_path)
        assert (tmp_path / "ours.txt").read_bytes() == (tmp_path / "ref.txt").read_bytes()


@pytest.mark.parametrize("at_in_quality", [False, True])
def test_fastq_expected_and_reference(tmp_path, at_in_quality):
    rng = np.random.default_rng(2)
    text, seqs = _fastq(rng, 40, at_in_quality)
    (tmp_path / "in.fq").write_text(text)
    _run(CONVERT, ["-q", "in.fq", "-o", "ours.txt"], tmp_path)
    if not at_in_quality:
        assert (tmp_path / "ours.txt").read_text() == "\n".join(seqs) + "\n"
    if (REF / "convert_int16").exists():      # including the reference's '@'-in-quality behaviour
        _run(REF / "convert_int16", ["-q", "in.fq", "-o", "ref.txt"], tmp_path)
        assert (tmp_path / "ours.txt").read_bytes() == (tmp_path / "ref.txt").read_bytes()


def _write_result(path, rng, nq, blocks, dtype):
    """blocks: list of (counts per device, extra_count).  Payload order (cal_cpu.c:363-401, thread.c:149-158):
    read bucket -> ref bucket (<= 100 queries) -> device -> [query][subject].  Returns the expected text order."""
    ndev = len(blocks[0][0])
    scores = {}   # (block, dev) -> [nq, count]
    with open(path, "wb") as f:
        for b, (counts, extra) in enumerate(blocks):
            for d, c in enumerate(counts):
                lo, hi = (-128, 127) if dtype == np.int8 else (-30000, 30000)
                scores[(b, d)] = rng.integers(lo, hi, size=(nq, c)).astype(dtype)
            for r0 in range(0, nq, 100):
                for d, c in enumerate(counts):
                    f.write(scores[(b, d)][r0:r0 + 100].tobytes())
    with open(str(path) + ".info", "wb") as f:
        f.write(struct.pack("<iiq", len(blocks), ndev, nq))
        for counts, extra in blocks:
            f.write(struct.pack("<%dq" % ndev, *counts) + struct.pack("<i", extra))
    out = []
    for q in range(nq):
        for b, (counts, extra) in enumerate(blocks):
            for d, c in enumerate(counts):
                keep = c - (extra if d == ndev - 1 else 0)
                out += [str(int(v)) for v in scores[(b, d)][q, :keep]]
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("nq,blocks,dtype", [
    (3, [([128], 0)], np.int16),
    (1, [([1000, 500], 12), ([64, 32], 0)], np.int16),
    (230, [([40, 24], 8), ([16, 16], 3), ([7, 9], 0)], np.int16),       # three ref buckets x three read buckets
    (101, [([33], 1)], np.int8),
    (5, [([10, 20, 30, 40], 5), ([4, 4, 4, 4], 0)], np.int8),
])
def test_result_conversion_expected_and_reference(tmp_path, nq, blocks, dtype):
    rng = np.random.default_rng(nq)
    expected = _write_result(tmp_path / "r.bin", rng, nq, blocks, dtype)
    b = "1" if dtype == np.int8 else "2"
    out = _run(CONVERT, ["-r", "r.bin", "-o", "ours.txt", "-b", b], tmp_path)
    assert (tmp_path / "ours.txt").read_text() == expected
    assert b"read_count[0][0] is %d" % blocks[0][0][0] in out
    ref = REF / ("convert_int8" if dtype == np.int8 else "convert_int16")
    if ref.exists():
        _run(ref, ["-r", "r.bin", "-o", "ref.txt"], tmp_path)
        assert (tmp_path / "ours.txt").read_bytes() == (tmp_path / "ref.txt").read_bytes()


def test_reads_the_reference_goldens(tmp_path):
    """The only checked-in result file of the reference (banded/BGSA_KNC/data, 2 devices) -- committed under
    tests/golden by make_golden.py together with the text the reference's convert made of it."""
    g = ROOT / "tests" / "golden"
    if not (g / "knc_result.bin").exists():
        pytest.skip("golden result file not committed")
    for name in ("knc_result.bin", "knc_result.bin.info"):
        (tmp_path / name).write_bytes((g / name).read_bytes())
    _run(CONVERT, ["-r", "knc_result.bin", "-o", "ours.txt"], tmp_path)
    assert (tmp_path / "ours.txt").read_bytes() == (g / "knc_convert_result.txt").read_bytes()
