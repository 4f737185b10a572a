"""CPU tier: host-side logic of the product -- the C ABI surface, kernel selection, the query
mask builder, the DP column functions (run on the host through tests/libhost_sim.so, same source
as the CUDA kernels) and the multi-rank sharding (gloo, world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import refutil as R

ROOT = Path(__file__).resolve().parent.parent


def _ensure_built():
    lib = ROOT / "bgsa_b200" / "libbgsa_b200.so"
    sim = ROOT / "tests" / "libhost_sim.so"
    if not lib.exists() or not sim.exists():
        subprocess.check_call(["make", "-s", "-j8", "-C", str(ROOT), "lib", "sim"])
    return lib, sim


@pytest.fixture(scope="module")
def sim():
    _, path = _ensure_built()
    lib = C.CDLL(str(path))
    lib.host_sim_align.restype = C.c_int
    lib.host_sim_align.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
    return lib


def run_sim(sim, algo, scheme, K, L, q, s, sign=-1):
    qc = np.ascontiguousarray(R.to_codes(q)); ss = np.ascontiguousarray(s)
    out = np.zeros((q.shape[0], s.shape[0]), np.int16)
    rc = sim.host_sim_align(algo, scheme, K, L, sign, qc.ctypes.data, q.shape[0], q.shape[1] - 1, ss.ctypes.data,
                            s.shape[0], s.shape[1] - 1, out.ctypes.data)
    return rc, out


def _instances(macro):
    text = (ROOT / "bgsa_b200" / "csrc" / "instances.h").read_text().replace("\\\n", " ")
    body = re.search(r"#define %s\(X\)(.*)" % macro, text).group(1)
    return [(int(a), int(b)) for a, b in re.findall(r"X\((\d+),\s*(\d+)\)", body)]


def test_c_abi_exports_every_declared_symbol():
    import bgsa_b200 as B
    _ensure_built()
    header = (ROOT / "include" / "bgsa_b200.h").read_text()
    declared = set(re.findall(r"\b(bgsa_[a-z_0-9]+)\s*\(", header))
    declared -= {"bgsa_status_t", "bgsa_algo_t", "bgsa_params_t", "bgsa_seq_t"}
    assert declared == set(B.EXPORTED_SYMBOLS)
    lib = B.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.bgsa_version()


def test_no_cpu_fallback_and_error_codes():
    import bgsa_b200 as B
    import torch
    p = B.Params.default(B.BITPAL_PACKED)
    assert (p.match, p.mismatch, p.gap, p.threshold, p.myers_sign) == (2, -3, -5, 31, -1)
    assert B.load().bgsa_result_size(B.BANDED_MYERS) == 1 and B.load().bgsa_result_size(B.MYERS_GLOBAL) == 2
    # unsupported requests are rejected before any CUDA call
    assert B.supported(B.Params.default(B.BITPAL_PACKED, match=7, mismatch=-1, gap=-3), 100, 100)     # not built in: instantiated at run time
    assert not B.supported(B.Params.default(B.BITPAL_PACKED, match=2, mismatch=3, gap=-5), 100, 100)  # not a scoring scheme
    assert not B.supported(B.Params.default(B.BITPAL_PACKED, match=60, mismatch=-1, gap=-30), 100, 100)   # delta range beyond the kernels
    assert not B.supported(B.Params.default(B.BANDED_MYERS, threshold=5), 100, 120)
    assert not B.supported(B.Params.default(B.BANDED_MYERS, threshold=40), 100, 100)
    assert not B.supported(B.Params.default(B.MYERS_GLOBAL), 40000, 100)
    assert B.supported(B.Params.default(B.MYERS_GLOBAL), 32768, 100)
    assert not B.supported(B.Params.default(B.MYERS_GLOBAL), 0, 100)
    if not torch.cuda.is_available():
        q = np.full((1, 11), 65, np.uint8); q[:, 10] = 10
        with pytest.raises(B.BgsaError) as ei:
            B.align_batch(p, q, np.repeat(q, 4, axis=0))
        assert ei.value.code == 3          # BGSA_ERR_CUDA: the product never falls back to the CPU


def test_kernel_selection_matches_baseline_configs():
    import bgsa_b200 as B
    assert B.kernel_name(B.Params.default(B.BITPAL_PACKED), 150, 150) == "align_kernel<BitpalPacked<2,-3,-5,K=5>,L=1>"
    assert B.kernel_name(B.Params.default(B.BITPAL_PACKED), 5000, 5000) == "align_kernel<BitpalPacked<2,-3,-5,K=10>,L=16>"
    assert B.kernel_name(B.Params.default(B.MYERS_SEMIGLOBAL), 1000, 1000) == "align_kernel<MyersAlgo<K=32,semiglobal>,L=1>"
    assert B.kernel_name(B.Params.default(B.MYERS_GLOBAL), 500, 500) == "align_kernel<MyersAlgo<K=16,global>,L=1>"
    assert B.kernel_name(B.Params.default(B.BANDED_MYERS, threshold=5), 100, 100) == "banded_kernel<u32>"
    assert B.kernel_name(B.Params.default(B.BANDED_MYERS, threshold=31), 100, 100) == "banded_kernel<u64>"


@pytest.mark.parametrize("layout_algo", [0, 2])
def test_host_pack_matches_numpy(layout_algo):
    """bgsa_pack_subjects_host (the host-thread front end of the batch entry, csrc/host_pack.cpp) writes the same tile
    layout as the device pack kernels: against the numpy packer, clean and dirty rows (N, lower case, arbitrary bytes,
    row ends that are not newlines), lengths around every vector / unit boundary, counts off the tile grid."""
    import bgsa_b200 as B
    _ensure_built()
    p = B.Params.default(layout_algo, threshold=5)
    layout = 1 if layout_algo == 2 else 0
    threads, isa = B.host_pack_info()
    assert threads >= 1 and isa in ("avx2", "scalar")
    for slen, n in [(150, 1000), (100, 4099), (1000, 130), (15, 77), (16, 64), (31, 33), (32, 5), (33, 64), (127, 97), (511, 65),
                    (63, 32), (64, 31), (65, 1), (5000, 40), (3, 50), (255, 200), (96, 3000)]:
        rng = np.random.default_rng(slen * 7 + n)
        for variant in ("clean", "dirty"):
            rows = R.random_rows(rng, n, slen, with_n=0.0 if variant == "clean" else 0.02)
            if variant == "dirty":
                junk = rng.random(rows[:, :slen].shape) < 0.01
                rows[:, :slen][junk] = rng.integers(0, 256, size=int(junk.sum()), dtype=np.uint8)
                rows[::7, slen] = 13
                rows[n // 2, :slen] = ord("A")
            # the rows sit at the very end of their allocation: an encoder that reads past the last row faults under ASan
            # and, here, would at least pick up the guard bytes
            buf = np.full(rows.size + 1, ord("T"), dtype=np.uint8)
            buf[: rows.size] = rows.reshape(-1)
            raw = B.pack_subjects_host(p, buf[: rows.size].reshape(n, slen + 1))
            codes, nm, flags = R.split_packed(raw, slen, n)
            ec, en, ef = R.numpy_pack(rows, layout)
            assert (codes == ec).all(), (slen, n, variant)
            assert ((flags != 0) == ef).all(), (slen, n, variant)
            assert (nm[ef] == en[ef]).all(), (slen, n, variant)


def test_host_pack_scalar_encoder_agrees(tmp_path):
    """The scalar twin of the AVX2 encoder (CPUs without AVX2; forced with BGSA_HOST_PACK_SCALAR=1) in a fresh process."""
    code = """
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import bgsa_b200 as B, refutil as R
assert B.host_pack_info()[1] == 'scalar'
rng = np.random.default_rng(5)
for algo in (0, 2):
    for slen, n in ((150, 333), (33, 70), (1000, 40)):
        rows = R.random_rows(rng, n, slen, with_n=0.02)
        raw = B.pack_subjects_host(B.Params.default(algo, threshold=5), rows)
        codes, nm, flags = R.split_packed(raw, slen, n)
        ec, en, ef = R.numpy_pack(rows, 1 if algo == 2 else 0)
        assert (codes == ec).all() and ((flags != 0) == ef).all() and (nm[ef] == en[ef]).all()
print('ok')
""" % (str(ROOT), str(ROOT / "tests"))
    env = dict(os.environ, BGSA_HOST_PACK_SCALAR="1", BGSA_HOST_THREADS="3")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr


def test_host_pool_many_small_batches_oversubscribed():
    """The worker pool under the conditions of one process per GPU on a shared host (bench.py --gpus N, aligner -g N):
    far more pool threads than cores, thousands of tiny parallel_for calls from several callers at once.  A batch lives
    on its caller's stack; a worker that still holds a pointer to it after the caller returned crashed rank 0 at N = 2
    (use after return) -- the pool now counts the holders."""
    code = """
import sys, threading, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import bgsa_b200 as B, refutil as R
assert B.host_pack_info()[0] == 48
rng = np.random.default_rng(11)
p = B.Params.default(0)
rows = R.random_rows(rng, 96, 40)
want = B.pack_subjects_host(p, rows).tobytes()
bad = []
def hammer():
    for _ in range(1500):
        if B.pack_subjects_host(p, rows).tobytes() != want:
            bad.append(1)
ts = [threading.Thread(target=hammer) for _ in range(3)]
[t.start() for t in ts]; [t.join() for t in ts]
assert not bad
print('ok')
""" % (str(ROOT), str(ROOT / "tests"))
    _ensure_built()
    env = dict(os.environ, BGSA_HOST_THREADS="48")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, (res.returncode, res.stderr[-2000:])


def test_rows_kernel_selection_needs_no_gpu():
    """Which short-read geometries get the one-kernel path (rows_kernel.cuh: ASCII tile by bulk copy, masks by byte value)
    is decided on the host: thread-per-subject instances up to 12 words, a tile that fits the stage, a row pitch whose 32
    lanes do not pile up on a few shared-memory banks.  Everything else is pack + align; banded has its own fused kernel."""
    import bgsa_b200 as B
    _ensure_built()
    my = B.Params.default(B.MYERS_GLOBAL)
    assert B.rows_kernel_name(my, 150, 150) == ("align_rows_kernel<MyersAlgo<K=5,global>>", True)
    assert B.rows_kernel_name(B.Params.default(B.BITPAL_PACKED), 150, 150) == ("align_rows_kernel<BitpalPacked<2,-3,-5,K=5>>", True)
    assert B.rows_kernel_name(B.Params.default(B.BITPAL_NONPACKED), 150, 150)[1]
    assert B.rows_kernel_name(B.Params.default(B.BITPAL_PACKED_SEMIGLOBAL), 100, 300)[1]
    assert B.rows_kernel_name(my, 384, 250)[1] and not B.rows_kernel_name(my, 385, 250)[1]          # 12 words
    name, fused = B.rows_kernel_name(my, 150, 1000)                                                 # tile too large for the stage
    assert not fused and name == "pack_stream_kernel + align_kernel<MyersAlgo<K=5,global>,L=1>"
    for slen, ok in ((127, False), (63, False), (191, False), (255, False), (100, True), (101, True), (151, True), (135, True), (250, True)):
        assert B.rows_kernel_name(my, 100, slen)[1] == ok, slen                                     # pitch = slen + 1
    assert not B.rows_kernel_name(B.Params.default(B.BITPAL_PACKED), 5000, 150)[1]                  # wavefront instance
    name, fused = B.rows_kernel_name(B.Params.default(B.BANDED_MYERS, threshold=5), 100, 100)
    assert fused and name.startswith("banded_kernel<u32> (fused")
    assert not B.rows_kernel_name(B.Params.default(B.BANDED_MYERS, threshold=5), 1000, 1000)[1]


def test_jit_precompile_needs_no_gpu(tmp_path, monkeypatch):
    """Scoring schemes outside the compiled list are instantiated by NVRTC from the kernel headers embedded in the library
    (csrc/jit.cu; the run-time counterpart of the reference's generator, Main.java:240-315).  The compile itself needs no
    GPU: the headers must stay NVRTC-clean, the instance lands in the disk cache, invalid schemes are refused with the
    reason, and without NVRTC an unlisted scheme is BGSA_ERR_UNSUPPORTED as before."""
    import bgsa_b200 as B
    _ensure_built()
    monkeypatch.setenv("BGSA_JIT_CACHE", str(tmp_path / "jit"))
    builtin = B.Params.default(B.BITPAL_PACKED)
    B.jit_precompile(builtin, 150, 150)                                   # built in: nothing to compile
    assert not (tmp_path / "jit").exists()
    cases = [(B.BITPAL_PACKED, (0, -1, -1), 150), (B.BITPAL_PACKED, (4, -6, -10), 100), (B.BITPAL_NONPACKED, (1, -1, -2), 150),
             (B.BITPAL_PACKED_SEMIGLOBAL, (5, -3, -4), 150), (B.BITPAL_PACKED, (1, -1, -2), 2000)]
    for algo, (M, I, G), ql in cases:
        p = B.Params.default(algo, match=M, mismatch=I, gap=G)
        assert B.supported(p, ql, ql)
        assert B.kernel_name(p, ql, ql).endswith("[NVRTC]")
        B.jit_precompile(p, ql, ql)
    files = sorted((tmp_path / "jit").glob("bgsa_*.bin"))
    assert len(files) == len(cases) and all(f.stat().st_size > 10_000 for f in files)
    stamp = [f.stat().st_mtime_ns for f in files]
    for algo, (M, I, G), ql in cases:                                     # second time: from the cache, nothing rewritten
        B.jit_precompile(B.Params.default(algo, match=M, mismatch=I, gap=G), ql, ql)
    assert stamp == [f.stat().st_mtime_ns for f in files]
    for M, I, G in ((2, 3, -5), (2, -3, 5), (-1, -2, -3), (3, 3, -1), (64, -1, -33)):
        p = B.Params.default(B.BITPAL_PACKED, match=M, mismatch=I, gap=G)
        assert not B.supported(p, 150, 150)
        with pytest.raises(B.BgsaError) as ei:
            B.jit_precompile(p, 150, 150)
        assert ei.value.code == 2 and "scoring scheme" in str(ei.value)
    code = """
import sys
sys.path.insert(0, %r)
import bgsa_b200 as B
p = B.Params.default(B.BITPAL_PACKED, match=4, mismatch=-6, gap=-10)
assert not B.supported(p, 150, 150)
assert B.supported(B.Params.default(B.BITPAL_PACKED), 150, 150)
print('ok')
""" % str(ROOT)
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BGSA_NO_JIT="1"), capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr


def test_host_pool_respects_cpu_affinity():
    """The worker pool is sized by the CPUs its creator may run on, not by the machine: under taskset / a cpuset / after
    bgsa_bind_thread_to_device the workers inherit the mask, and more workers than allowed CPUs only slow each other down."""
    if not hasattr(os, "sched_setaffinity") or len(os.sched_getaffinity(0)) < 2:
        pytest.skip("needs at least two allowed CPUs")
    _ensure_built()
    code = """
import os, sys
sys.path.insert(0, %r)
cpus = sorted(os.sched_getaffinity(0))[:2]
os.sched_setaffinity(0, set(cpus))
import bgsa_b200 as B
print('threads', B.host_pack_info()[0])
""" % str(ROOT)
    env = {k: v for k, v in os.environ.items() if k not in ("BGSA_HOST_THREADS", "BGSA_HOST_GPUS", "LOCAL_WORLD_SIZE")}
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    assert int(res.stdout.split()[-1]) in (1, 2), res.stdout


def test_sass_carry_chains_and_budget():
    """Build-time guard for the hardware carry chains (bgsa_common.cuh add_chain: consecutive add.cc / addc.cc asm
    statements rely on nothing clobbering CC.CF in between): in the SASS of the thread-per-subject kernels every
    K-word add must be exactly one IADD3 that starts a carry chain plus K-1 links (IADD3.X, or IMAD.X for a dead
    carry-out) -- Myers: one chain per column, BitPAl packed (2,-3,-5): one per high class (NH = 5).  A toolkit that
    breaks or pads the chains changes these counts.  Also pins the ALU-pipe instruction budget per word-column that
    DESIGN.md and bench.py's roofline.frac_sass quote (tools/sass_budget.py)."""
    _ensure_built()
    sys.path.insert(0, str(ROOT / "tools"))
    import sass_budget as SB
    funcs = SB.sass_functions(ROOT / "bgsa_b200" / "libbgsa_b200.so")
    expect = {"C2_bitpal_packed_K5": (5, 5, 63.0), "C4_myers_semi_K32": (1, 32, 10.3), "myers150_K5": (1, 5, 10.2),
              "bitpal_nonpacked_150": (5, 5, 157.0),
              # the rows kernels (ASCII in, masks by byte value): the recurrence alone, word 0's shifts on the FMA pipe
              "C2_rows_bitpal_packed_K5": (5, 5, 62.5), "myers150_rows_K5": (1, 5, 9.8), "C2np_rows_K5": (5, 5, 156.0)}
    for name, (chains, K, max_alu_per_word) in expect.items():
        r = SB.analyse(name, SB.KERNELS[name], funcs, None)
        assert "error" not in r, (name, r)
        assert r["carry_chain_starts_per_column"] == chains, (name, r)
        assert r["carry_chain_links_per_column"] == chains * (K - 1), (name, r)
        assert r["alu_per_word_column"] <= max_alu_per_word, (name, r["alu_per_word_column"])
    r = SB.analyse("C3_banded_fused", SB.KERNELS["C3_banded_fused"], funcs, None)
    assert r["alu_per_column"] <= 13.5, r          # 13 per band row: no predicated-off N-path instructions in the common path


@pytest.mark.parametrize("mode", [0, 1])
def test_myers_columns_on_host(sim, mode):
    rng = np.random.default_rng(100 + mode)
    for K, L in _instances("BGSA_MYERS_INSTANCES"):
        cap = 32 * K * L
        for ql in {cap, max(1, cap - 33)}:
            if ql > 2600:
                ql = int(rng.integers(cap // 2 + 1, min(cap, 2600) + 1)) if cap // 2 < 2600 else 0
            if ql < 1:
                continue
            sl = int(rng.integers(max(1, ql // 3), ql + 40))
            q = R.random_rows(rng, 1, ql, with_n=0.02)
            s = R.random_rows(rng, 4, sl, with_n=0.02)
            s[0, : min(ql, sl)] = q[0, : min(ql, sl)]
            rc, got = run_sim(sim, mode, 0, K, L, q, s)
            assert rc == 0
            assert (got == R.oracle_batch(mode, q, s)).all(), (K, L, ql, sl)
    # generator -m 1: +distance (Main.java:127-141)
    q = R.random_rows(rng, 1, 90); s = R.random_rows(rng, 5, 80)
    rc, got = run_sim(sim, mode, 0, 3, 1, q, s, sign=1)
    assert rc == 0 and (got == -R.oracle_batch(mode, q, s)).all()


@pytest.mark.parametrize("scheme,mig", [(0, (2, -3, -5)), (1, (1, -1, -1)), (2, (1, -3, -2))])
@pytest.mark.parametrize("packed", [True, False])
def test_bitpal_columns_on_host(sim, scheme, mig, packed):
    rng = np.random.default_rng(200 + scheme)
    M, I, G = mig
    inst = _instances("BGSA_BITPAL_PACKED_INSTANCES" if packed else "BGSA_BITPAL_NONPACKED_INSTANCES")
    for K, L in inst:
        cap = min(32 * K * L, 1200 if packed else 700)
        for ql in {cap, max(1, cap - 1)}:
            sl = int(rng.integers(max(1, ql // 2), ql + 20))
            q = R.random_rows(rng, 1, ql, with_n=0.02)
            s = R.random_rows(rng, 3, sl, with_n=0.02)
            s[0, : min(ql, sl)] = q[0, : min(ql, sl)]
            rc, got = run_sim(sim, 3 if packed else 4, scheme, K, L, q, s)
            assert rc == 0
            assert (got == R.oracle_batch(R.ALGO_BITPAL_PACKED, q, s, M=M, I=I, G=G)).all(), (K, L, ql, sl)


@pytest.mark.parametrize("scheme,mig", [(0, (2, -3, -5)), (1, (1, -1, -1)), (2, (1, -3, -2))])
def test_bitpal_semiglobal_columns_on_host(sim, scheme, mig):
    """Semi-global BitPAl (whole query inside the subject): every (K, L) instance on the host against the
    restated generator emission AND plain DP.  Scheme 2 has -G in a high class (boundary enters a carry chain)."""
    rng = np.random.default_rng(300 + scheme)
    M, I, G = mig
    for K, L in _instances("BGSA_BITPAL_PACKED_INSTANCES"):
        cap = min(32 * K * L, 900)
        for ql in {cap, max(1, cap - 33), max(1, cap // 2 + 1)}:
            sl = int(rng.integers(max(1, ql // 2), 2 * ql + 20))
            q = R.random_rows(rng, 1, ql, with_n=0.02)
            s = R.random_rows(rng, 3, sl, with_n=0.02)
            if sl > ql + 4:
                s[0, 3:3 + ql] = q[0, :ql]                      # the query planted inside a subject
            s[1, : min(ql, sl)] = q[0, : min(ql, sl)]
            rc, got = run_sim(sim, 5, scheme, K, L, q, s)
            assert rc == 0
            assert (got == R.oracle_batch(R.ALGO_BITPAL_SEMI, q, s, M=M, I=I, G=G)).all(), (K, L, ql, sl)
            assert (got == R.dp_scores("nw_semi", q, s, M=M, I=I, G=G).astype(np.int16)).all(), (K, L, ql, sl)


def test_other_scoring_schemes_match_dp(tmp_path):
    """`make SCHEMES=...` = re-running the reference's generator for another (M, I, G): the same sources
    instantiated for five more schemes (common factor 2, zero match score, -G in a high class, ...) on the host
    simulator against plain DP -- packed global, packed semi-global and non-packed."""
    import subprocess
    schemes = [(3, -2, -4), (4, -6, -10), (1, -1, -2), (5, -3, -4), (0, -1, -1)]
    hdr = tmp_path / "schemes.h"
    hdr.write_text("#define BGSA_SCHEMES(X) " + " ".join(f"X({i}, {m}, {x}, {g})" for i, (m, x, g) in enumerate(schemes)) + "\n")
    so = tmp_path / "libhost_sim_schemes.so"
    subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I", str(ROOT / "bgsa_b200" / "csrc"),
                           "-include", str(hdr), "-o", str(so), str(ROOT / "tests" / "host_sim.cu")],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lib = C.CDLL(str(so))
    lib.host_sim_align.restype = C.c_int
    lib.host_sim_align.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
    rng = np.random.default_rng(77)
    for sid, (M, I, G) in enumerate(schemes):
        for K, L, algo, kind in [(5, 1, 3, "nw"), (6, 2, 3, "nw"), (10, 2, 5, "nw_semi"), (3, 1, 5, "nw_semi"), (2, 2, 4, "nw")]:
            ql = min(32 * K * L, 260) - int(rng.integers(0, 20))
            sl = int(rng.integers(ql // 2, ql + 60))
            q = R.random_rows(rng, 1, ql, with_n=0.02)
            s = R.random_rows(rng, 4, sl, with_n=0.02)
            s[0, : min(ql, sl)] = q[0, : min(ql, sl)]
            if sl > ql + 5:
                s[1, 3:3 + ql] = q[0, :ql]
            rc, got = run_sim(lib, algo, sid, K, L, q, s)
            assert rc == 0
            assert (got == R.dp_scores(kind, q, s, M=M, I=I, G=G).astype(np.int16)).all(), ((M, I, G), K, L, algo)


def test_shard_counts():
    from bgsa_b200.sharding import shard_counts, shard_range
    assert shard_counts(1_000_000, 8) == [124992] * 7 + [125056]
    assert shard_counts(100, 8) == [0] * 7 + [100]
    assert shard_counts(0, 2) == [0, 0]
    for total in (1, 31, 32, 33, 1000, 12345):
        for world in (1, 2, 3, 8):
            counts = shard_counts(total, world)
            assert sum(counts) == total and all(c % 32 == 0 for c in counts[:-1])
            assert [shard_range(total, r, world)[1] for r in range(world)] == counts
            assert shard_range(total, world - 1, world)[0] + counts[-1] == total


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests"); sys.path.insert(0, {root!r} + "/tools")
import numpy as np, torch, torch.distributed as dist
import refutil as R, synth
from bgsa_b200.sharding import shard_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
q, s = synth.make("C2", 1000)
first, count = shard_range(s.shape[0], rank, world)
# each rank scores its own contiguous range (here with the oracle standing in for the GPU) ...
mine = R.oracle_batch(R.ALGO_BITPAL_PACKED, q, s[first:first + count])
# ... and the scores are gathered to rank 0 in device-major order, no collective on the data path
parts = [None] * world
dist.gather_object((first, mine), parts if rank == 0 else None, dst=0)
t = torch.tensor([float(count)]); dist.all_reduce(t)
assert int(t.item()) == s.shape[0]
if rank == 0:
    full = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])], axis=1)
    assert (full == R.oracle_batch(R.ALGO_BITPAL_PACKED, q, s)).all()
    print("OK")
dist.destroy_process_group()
"""


def test_two_rank_sharding_gloo(tmp_path):
    import socket
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT), port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0]


# ---- pack: SWAR encoder of csrc/pack.cuh::encode_piece, modelled in Python, exhaustive per byte ------
def _swar_word(w):
    M = 0xFFFFFFFF
    s1, l2, l3, l4 = w >> 1, (w << 2) & M, (w << 3) & M, (w << 4) & M
    x = (w ^ s1) & 0x06060606
    v2 = (w ^ ((~l2 & M) | l3)) & M
    g = (w ^ l4) & v2
    bad = ((w ^ 0x40404040) & 0xE8E8E8E8) | (~g & 0x10101010)
    K = (1 << 23) | (1 << 17) | (1 << 11) | (1 << 5)
    KL = (1 << 27) | (1 << 20) | (1 << 13) | (1 << 6)
    return bad, ((x * K) & M) >> 24, (((x & 0x02020202) * KL) & M) >> 28, (((x & 0x04040404) * (KL >> 1)) & M) >> 28


def test_pack_swar_model():
    code = {0x41: 0, 0x43: 1, 0x47: 2, 0x54: 3}
    rng = np.random.default_rng(3)
    for pos in range(4):
        for c in range(256):
            for trial in range(12):
                bs = [int(b) for b in (rng.choice([0x41, 0x43, 0x47, 0x54], 4) if trial % 3 else rng.integers(0, 256, 4))]
                bs[pos] = c
                w = sum(b << (8 * i) for i, b in enumerate(bs))
                bad, c8, lo4, hi4 = _swar_word(w)
                for i, b in enumerate(bs):      # the validity test is exact per byte
                    assert (((bad >> (8 * i)) & 0xFF) == 0) == (b in code)
                if all(b in code for b in bs):  # and the gathered codes / planes are right when it passes
                    assert c8 == sum(code[b] << (2 * i) for i, b in enumerate(bs))
                    assert lo4 == sum((code[b] & 1) << i for i, b in enumerate(bs))
                    assert hi4 == sum((code[b] >> 1) << i for i, b in enumerate(bs))
