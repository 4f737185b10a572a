"""GPU tier: the drop-in boundary end to end.
  * our C `aligner` (same -q/-d/-f command line) writes a result file + .info that the REFERENCE's
    own `convert -r` (oracle/_ref/convert_int16|int8, compiled from the unmodified reference) reads,
    and the bytes equal the reference aligner's output;
  * the reference's UNMODIFIED host pipeline linked against our align_core shim (oracle/_ref/dropin_*)
    produces the same files with the DP on the GPU.
Nothing here reads /root/reference: the reference binaries travel prebuilt in oracle/_ref/."""
import hashlib
import os
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

import refutil as R
import synth

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
ALIGNER = ROOT / "bgsa_b200" / "aligner"
REF = ROOT / "oracle" / "_ref"


def md5(path):
    return hashlib.md5(Path(path).read_bytes()).hexdigest()


def run(cmd, cwd):
    res = subprocess.run([str(c) for c in cmd], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout
    return res.stdout


def need(path):
    if not Path(path).exists():
        pytest.skip(f"{path} not present (built only where the reference sources are available)")


@pytest.fixture(scope="module")
def sample(tmp_path_factory):
    d = tmp_path_factory.mktemp("sample")
    q, s = R.sample_data()
    R.write_rows(d / "query.txt", q)
    R.write_rows(d / "subject.txt", s)
    return d


def test_aligner_myers_sample_files_and_reference_convert(sample):
    assert ALIGNER.exists(), "bgsa_b200/aligner not built (make tools)"
    out = run([ALIGNER, "-q", "query.txt", "-d", "subject.txt", "-f", "r_myers.bin"], sample)
    assert "cal GCUPS is" in out and "Total GCUPS is" in out and "score is 0, -1, -1" in out
    # byte-identical to the reference build on the same inputs (SURVEY.md section 4 / BASELINE.md section 2)
    assert md5(sample / "r_myers.bin") == "7253c1f2a6423aaa3e29577acc137302"
    assert md5(sample / "r_myers.bin.info") == "210a66912c2846fdc6d3e64fa8fbe61f"
    need(REF / "convert_int16")
    run([REF / "convert_int16", "-r", "r_myers.bin", "-o", "r_myers.txt"], sample)
    assert md5(sample / "r_myers.txt") == "862d379d256f1b7c7efe0d0505f08958"
    assert (sample / "r_myers.txt").read_text().split()[:5] == ["-269", "-275", "-263", "-277", "-289"]


def test_aligner_bitpal_and_semiglobal_sample(sample, golden_dir):
    run([ALIGNER, "-a", "bitpal", "-q", "query.txt", "-d", "subject.txt", "-f", "r_bitpal.bin"], sample)
    assert md5(sample / "r_bitpal.bin").startswith("6dcc0519e8")
    run([ALIGNER, "-a", "bitpal-nonpacked", "-q", "query.txt", "-d", "subject.txt", "-f", "r_bitpal_np.bin"], sample)
    assert md5(sample / "r_bitpal_np.bin").startswith("6dcc0519e8")
    run([ALIGNER, "-a", "semiglobal", "-q", "query.txt", "-d", "subject.txt", "-f", "r_semi.bin"], sample)
    got = np.fromfile(sample / "r_semi.bin", dtype=np.int16).reshape(3, 128)
    assert (got == np.load(golden_dir / "golden_semiglobal_knc.npz")["scores"]).all()
    if (REF / "convert_int16").exists():
        run([REF / "convert_int16", "-r", "r_bitpal.bin", "-o", "r_bitpal.txt"], sample)
        assert md5(sample / "r_bitpal.txt").startswith("a2e99d94f83e")
        run([REF / "convert_int16", "-r", "r_semi.bin", "-o", "r_semi.txt"], sample)
        assert md5(sample / "r_semi.txt") == "41696989c6d7e5f58897d81d154c47ea"     # the checked-in convert_result.txt


def test_aligner_any_scoring_scheme(sample, tmp_path, monkeypatch):
    """`aligner -a bitpal -M -I -G` with a scheme that is not built into the library: the reference would re-run its
    generator and recompile (Main.java:240-315); here the instance is made at run time (csrc/jit.cu) -- same CLI, scores
    equal to plain DP, result file readable by the reference's convert."""
    monkeypatch.setenv("BGSA_JIT_CACHE", str(tmp_path / "jit"))
    q, s = R.sample_data()
    for name, algo, kind, (M, I, G) in (("bitpal", R.ALGO_BITPAL_PACKED, "nw", (4, -6, -10)),
                                        ("bitpal-nonpacked", R.ALGO_BITPAL_PACKED, "nw", (1, -1, -2)),
                                        ("bitpal-semiglobal", R.ALGO_BITPAL_SEMI, "nw_semi", (3, -2, -4))):
        out = f"r_{name}_{M}.bin"
        run([ALIGNER, "-a", name, "-M", str(M), "-I", str(I), "-G", str(G), "-q", "query.txt", "-d", "subject.txt", "-f", out], sample)
        got = np.fromfile(sample / out, dtype=np.int16).reshape(3, 128)
        assert (got == R.oracle_batch(algo, q, s, M=M, I=I, G=G)).all(), (name, M, I, G)
        assert (got[:, :4] == R.dp_scores(kind, q, s[:4], M=M, I=I, G=G).astype(np.int16)).all()
    assert len(list((tmp_path / "jit").glob("bgsa_*.bin"))) == 3
    if (REF / "convert_int16").exists():
        run([REF / "convert_int16", "-r", "r_bitpal_4.bin", "-o", "r_bitpal_4.txt"], sample)
        assert len((sample / "r_bitpal_4.txt").read_text().split()) == 3 * 128


def test_aligner_query_file_without_final_newline_and_m1(sample):
    q, _ = R.sample_data()
    R.write_rows(sample / "query_nonl.txt", q, final_newline=False)
    run([ALIGNER, "-q", "query_nonl.txt", "-d", "subject.txt", "-f", "r_nonl.bin"], sample)
    assert md5(sample / "r_nonl.bin") == "7253c1f2a6423aaa3e29577acc137302"
    run([ALIGNER, "-m", "1", "-q", "query.txt", "-d", "subject.txt", "-f", "r_pos.bin"], sample)
    a = np.fromfile(sample / "r_pos.bin", dtype=np.int16); b = np.fromfile(sample / "r_myers.bin", dtype=np.int16)
    assert (a == -b).all()


def test_aligner_banded_c3_slice_vs_reference_aligner(tmp_path):
    q, s = synth.make("C3", 50_000)
    R.write_rows(tmp_path / "q.txt", q); R.write_rows(tmp_path / "s.txt", s)
    run([ALIGNER, "-a", "banded", "-k", "5", "-q", "q.txt", "-d", "s.txt", "-f", "ours.bin"], tmp_path)
    ours = np.fromfile(tmp_path / "ours.bin", dtype=np.int8)
    assert (ours[None, :] == R.oracle_batch(R.ALGO_BANDED, q, s, e=5)).all()
    if (REF / "aligner_banded_cpu").exists():
        run([REF / "aligner_banded_cpu", "-k", "5", "-q", "q.txt", "-d", "s.txt", "-f", "ref.bin"], tmp_path)
        assert md5(tmp_path / "ours.bin") == md5(tmp_path / "ref.bin")
        assert md5(tmp_path / "ours.bin.info") == md5(tmp_path / "ref.bin.info")
        run([REF / "convert_int8", "-r", "ours.bin", "-o", "ours.txt"], tmp_path)
        run([REF / "convert_int8", "-r", "ref.bin", "-o", "ref.txt"], tmp_path)
        assert md5(tmp_path / "ours.txt") == md5(tmp_path / "ref.txt")


def test_aligner_default_result_path(sample, tmp_path):
    """Without -f the reference writes data/result.txt (+ .info) and creates ./data itself (original/BGSA_CPU/main.c:39,
    cal_cpu.c:198); so do we, and the reference aligner run the same way leaves identical files."""
    out = run([ALIGNER, "-q", sample / "query.txt", "-d", sample / "subject.txt"], tmp_path)
    assert "cal GCUPS is" in out
    assert md5(tmp_path / "data" / "result.txt") == "7253c1f2a6423aaa3e29577acc137302"
    assert md5(tmp_path / "data" / "result.txt.info") == "210a66912c2846fdc6d3e64fa8fbe61f"
    if (REF / "aligner_myers_cpu").exists():
        ref_dir = tmp_path / "ref"
        ref_dir.mkdir()
        run([REF / "aligner_myers_cpu", "-q", sample / "query.txt", "-d", sample / "subject.txt"], ref_dir)
        assert md5(ref_dir / "data" / "result.txt") == md5(tmp_path / "data" / "result.txt")
        assert md5(ref_dir / "data" / "result.txt.info") == md5(tmp_path / "data" / "result.txt.info")


def test_aligner_multi_bucket_multi_query_vs_reference_aligner(tmp_path):
    """> READ_BUCKET_SIZE (114857600 B) of subjects => 2 read buckets; 3 queries => [query][subject] per bucket."""
    need(REF / "aligner_myers_cpu")
    q, _ = synth.make("C2", 1)
    rng = np.random.default_rng(5)
    q3 = np.concatenate([q, R.random_rows(rng, 2, 150)])
    _, s = synth.make("C2", 800_000)
    R.write_rows(tmp_path / "q.txt", q3); R.write_rows(tmp_path / "s.txt", s)
    run([ALIGNER, "-q", "q.txt", "-d", "s.txt", "-f", "ours.bin"], tmp_path)
    run([REF / "aligner_myers_cpu", "-q", "q.txt", "-d", "s.txt", "-f", "ref.bin"], tmp_path)
    assert md5(tmp_path / "ours.bin") == md5(tmp_path / "ref.bin")
    assert md5(tmp_path / "ours.bin.info") == md5(tmp_path / "ref.bin.info")
    nblocks, ndev, nq = struct.unpack("<iiq", (tmp_path / "ours.bin.info").read_bytes()[:16])
    assert (nblocks, ndev, nq) == (2, 1, 3)
    run([REF / "convert_int16", "-r", "ours.bin", "-o", "ours.txt"], tmp_path)
    txt = np.loadtxt(tmp_path / "ours.txt", dtype=np.int64).reshape(3, -1)
    idx = rng.choice(s.shape[0], 300, replace=False)
    assert (txt[:, idx] == R.oracle_batch(R.ALGO_MYERS_GLOBAL, q3, np.ascontiguousarray(s[idx]))).all()


def test_aligner_many_items_in_flight_order(tmp_path):
    """5 read buckets x 3 ref buckets (230 queries) = 15 pipelined items, two in flight: the payload must still be
    read bucket -> ref bucket -> [query][subject] (cal_cpu.c:363-401), which convert -r undoes."""
    convert = ROOT / "bgsa_b200" / "convert"
    rng = np.random.default_rng(9)
    q = R.random_rows(rng, 230, 60, with_n=0.01)
    s = np.concatenate([R.mutate_rows(rng, q[7, :60], 2000, 12), R.random_rows(rng, 2611, 60, with_n=0.01)])
    R.write_rows(tmp_path / "q.txt", q); R.write_rows(tmp_path / "s.txt", s)
    env = dict(os.environ, BGSA_READ_BUCKET_SIZE=str(61 * 1000))          # 1000 rows per read bucket
    for algo, oalgo, b in (("bitpal", R.ALGO_BITPAL_PACKED, "2"), ("myers", R.ALGO_MYERS_GLOBAL, "2")):
        res = subprocess.run([str(ALIGNER), "-a", algo, "-q", "q.txt", "-d", "s.txt", "-f", "many.bin"], cwd=tmp_path, env=env,
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert res.returncode == 0, res.stdout
        nblocks, ndev, nq = struct.unpack("<iiq", (tmp_path / "many.bin.info").read_bytes()[:16])
        assert (nblocks, ndev, nq) == (5, 1, 230)
        run([convert, "-r", "many.bin", "-o", "many.txt", "-b", b], tmp_path)
        txt = np.loadtxt(tmp_path / "many.txt", dtype=np.int64).reshape(230, -1)
        assert (txt == R.oracle_batch(oalgo, q, s)).all(), algo
        if (REF / "convert_int16").exists():
            run([REF / "convert_int16", "-r", "many.bin", "-o", "many_ref.txt"], tmp_path)
            assert md5(tmp_path / "many.txt") == md5(tmp_path / "many_ref.txt")


def test_aligner_two_gpus_device_major_file(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    need(REF / "convert_int16")
    rng = np.random.default_rng(6)
    q = R.random_rows(rng, 3, 150); s = R.random_rows(rng, 10_007, 150)
    R.write_rows(tmp_path / "q.txt", q); R.write_rows(tmp_path / "s.txt", s)
    run([ALIGNER, "-a", "bitpal", "-g", "2", "-q", "q.txt", "-d", "s.txt", "-f", "two.bin"], tmp_path)
    info = (tmp_path / "two.bin.info").read_bytes()
    assert struct.unpack("<iiq", info[:16]) == (1, 2, 3)
    c0, c1, extra = struct.unpack("<qqi", info[16:36])
    assert c0 % 32 == 0 and c0 + c1 == 10_007 and extra == 0
    run([REF / "convert_int16", "-r", "two.bin", "-o", "two.txt"], tmp_path)
    txt = np.loadtxt(tmp_path / "two.txt", dtype=np.int64).reshape(3, -1)
    assert (txt == R.oracle_batch(R.ALGO_BITPAL_PACKED, q, s)).all()
    # long rows on the second device too (the pack kernel opts in to > 48 KB of shared memory per device)
    q5, s5 = synth.make("C5", 300)
    R.write_rows(tmp_path / "q5.txt", q5); R.write_rows(tmp_path / "s5.txt", s5)
    run([ALIGNER, "-a", "bitpal", "-g", "2", "-q", "q5.txt", "-d", "s5.txt", "-f", "two5.bin"], tmp_path)
    got = np.fromfile(tmp_path / "two5.bin", dtype=np.int16)
    assert (got[None, :] == R.oracle_batch(R.ALGO_BITPAL_PACKED, q5, s5)).all()


# ---- the reference's own, unmodified host pipeline on top of our align_core shims -----------------
DROPIN = [("myers_cpu", [], "7253c1f2a6423aaa3e29577acc137302", None),
          ("myers_sse", [], "7253c1f2a6423aaa3e29577acc137302", "sse4_1"),
          ("bitpal_avx2", [], "6dcc0519e8", "avx2"),
          ("bitpal_avx512", [], "6dcc0519e8", "avx512f")]


@pytest.mark.parametrize("variant,extra,digest,flag", DROPIN)
def test_reference_pipeline_with_our_align_core(sample, variant, extra, digest, flag):
    need(REF / f"dropin_{variant}")
    if flag and flag not in R.cpu_flags():
        pytest.skip(f"host CPU lacks {flag}")
    env = dict(os.environ, OMP_NUM_THREADS="4")
    res = subprocess.run([str(REF / f"dropin_{variant}"), "-N", "4", "-q", "query.txt", "-d", "subject.txt", "-f", f"d_{variant}.bin"] + extra,
                         cwd=sample, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout
    assert md5(sample / f"d_{variant}.bin").startswith(digest)


def test_reference_pipeline_semiglobal_and_banded_dropin(sample, tmp_path, golden_dir):
    need(REF / "dropin_semiglobal_cpu")
    env = dict(os.environ, OMP_NUM_THREADS="4")
    subprocess.run([str(REF / "dropin_semiglobal_cpu"), "-N", "4", "-q", "query.txt", "-d", "subject.txt", "-f", "d_semi.bin"],
                   cwd=sample, env=env, check=True, stdout=subprocess.DEVNULL, timeout=600)
    got = np.fromfile(sample / "d_semi.bin", dtype=np.int16).reshape(3, 128)
    assert (got == np.load(golden_dir / "golden_semiglobal_knc.npz")["scores"]).all()
    q, s = synth.make("C3", 4_000)
    R.write_rows(tmp_path / "q.txt", q); R.write_rows(tmp_path / "s.txt", s)
    subprocess.run([str(REF / "dropin_banded_cpu"), "-N", "4", "-k", "5", "-q", "q.txt", "-d", "s.txt", "-f", "d_banded.bin"],
                   cwd=tmp_path, env=env, check=True, stdout=subprocess.DEVNULL, timeout=600)
    got = np.fromfile(tmp_path / "d_banded.bin", dtype=np.int8)
    assert (got[None, :] == R.oracle_batch(R.ALGO_BANDED, q, s, e=5)).all()


def test_fasta_to_scores_with_our_convert_tool(tmp_path):
    """The whole tool chain of the reference README: convert -f / -q, aligner, convert -r -- all ours."""
    convert = ROOT / "bgsa_b200" / "convert"
    assert convert.exists(), "bgsa_b200/convert not built (make tools)"
    rng = np.random.default_rng(8)
    q = R.random_rows(rng, 2, 120)
    s = np.concatenate([R.mutate_rows(rng, q[0, :120], 300, 10), R.random_rows(rng, 212, 120, with_n=0.01)])
    fasta = "".join(f">s{i} x\n{bytes(r[:70]).decode()}\n{bytes(r[70:120]).decode()}\n" for i, r in enumerate(s))
    fastq = "".join(f"@q{i}\n{bytes(r[:120]).decode()}\n+\n{'I' * 120}\n" for i, r in enumerate(q))
    (tmp_path / "s.fa").write_text(fasta); (tmp_path / "q.fq").write_text(fastq)
    run([convert, "-f", "s.fa", "-o", "s.txt"], tmp_path)
    run([convert, "-q", "q.fq", "-o", "q.txt"], tmp_path)
    assert (tmp_path / "s.txt").read_bytes() == s.tobytes() and (tmp_path / "q.txt").read_bytes() == q.tobytes()
    for algo, oalgo, bytes_per, extra in (("bitpal", R.ALGO_BITPAL_PACKED, "2", []), ("banded", R.ALGO_BANDED, "1", ["-k", "7"])):
        run([ALIGNER, "-a", algo, "-q", "q.txt", "-d", "s.txt", "-f", "r.bin", "-g", "1"] + extra, tmp_path)
        run([convert, "-r", "r.bin", "-o", "r.txt", "-b", bytes_per], tmp_path)
        got = np.array((tmp_path / "r.txt").read_text().split(), dtype=np.int64).reshape(2, -1)
        assert (got == R.oracle_batch(oalgo, q, s, e=7)).all()
