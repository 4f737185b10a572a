"""GPU tier (-m gpu): the CUDA path, called through the C ABI (bgsa_b200 ctypes binding of
include/bgsa_b200.h), must be bit-identical to the oracle / the reference vectors.  Nothing here
reads /root/reference.  Integer results: the bar is exact equality everywhere."""
import numpy as np
import pytest

import refutil as R
import synth

pytestmark = pytest.mark.gpu

ORACLE_ALGO = {0: R.ALGO_MYERS_GLOBAL, 1: R.ALGO_MYERS_SEMIGLOBAL, 2: R.ALGO_BANDED, 3: R.ALGO_BITPAL_PACKED,
               4: R.ALGO_BITPAL_PACKED}


@pytest.fixture(scope="module")
def B():
    import torch
    assert torch.cuda.is_available(), "GPU tier needs a CUDA device"
    import bgsa_b200 as B
    B.load()
    return B


@pytest.fixture(scope="module")
def vec(golden_dir):
    return np.load(golden_dir / "ref_vectors.npz")


def gpu(B, algo, q, s, **kw):
    before = B.launch_count()
    out = B.align_batch(B.Params.default(algo, **kw), q, s)
    assert B.launch_count() > before or s.shape[0] == 0     # our kernels really ran
    return out


def expect(algo, q, s, **kw):
    return R.oracle_batch(ORACLE_ALGO[algo], q, s, M=kw.get("match", 2), I=kw.get("mismatch", -3), G=kw.get("gap", -5),
                          e=kw.get("threshold", 5))


# ---- C1: the reference's own fixture, and its only checked-in golden ------------------------------
def test_c1_sample_data_md5(B, vec):
    import hashlib
    q, s = R.sample_data()
    got = gpu(B, B.MYERS_GLOBAL, q, s)
    assert hashlib.md5(got.tobytes()).hexdigest() == "7253c1f2a6423aaa3e29577acc137302"   # original/BGSA_CPU result.bin
    assert (got == vec["sample_myers_cpu"]).all()
    bp = gpu(B, B.BITPAL_PACKED, q, s)
    assert (bp == vec["sample_bitpal_avx512"]).all()
    assert (gpu(B, B.BITPAL_NONPACKED, q, s) == bp).all()
    assert (gpu(B, B.BANDED_MYERS, q, s, threshold=31) == vec["sample_banded_k31"]).all()


def test_semiglobal_checked_in_golden(B, golden_dir):
    g = np.load(golden_dir / "golden_semiglobal_knc.npz")
    q, s = R.sample_data()
    assert (gpu(B, B.MYERS_SEMIGLOBAL, q, s) == g["scores"]).all()


# ---- slices of C2..C5 against vectors produced by the unmodified reference ------------------------
@pytest.mark.parametrize("name,algo,key,kw", [
    ("C2", 3, "C2_ref", {}), ("C2", 4, "C2_ref", {}), ("C2", 0, "C2_myers_ref", {}),
    ("C3", 2, "C3_ref", {"threshold": 5}),
    ("C4", 1, "C4_ref_restated", {}), ("C4", 0, "C4_myers_ref", {}),
    ("C5", 3, "C5_ref", {}),
])
def test_config_slices_vs_reference_vectors(B, vec, name, algo, key, kw):
    got = gpu(B, algo, vec[f"{name}_query"], vec[f"{name}_subjects"], **kw)
    assert (got == vec[key]).all()


@pytest.mark.parametrize("i", range(5))
def test_ragged_and_n_vs_reference_vectors(B, vec, i):
    q, s = vec[f"rag{i}_query"], vec[f"rag{i}_subjects"]
    assert (gpu(B, B.MYERS_GLOBAL, q, s) == vec[f"rag{i}_myers"]).all()
    assert (gpu(B, B.BITPAL_PACKED, q, s) == vec[f"rag{i}_bitpal"]).all()
    assert (gpu(B, B.BITPAL_NONPACKED, q, s) == vec[f"rag{i}_bitpal"]).all()


# ---- every kernel geometry (K, L) against the oracle ----------------------------------------------
LENGTHS = [(1, 1, 3), (31, 40, 33), (32, 32, 64), (33, 20, 65), (96, 100, 130), (150, 150, 700), (160, 150, 97),
           (161, 300, 60), (256, 256, 100), (500, 480, 200), (640, 100, 70), (1000, 1000, 90), (1024, 64, 40),
           (1025, 200, 40), (2048, 300, 36), (2100, 500, 35), (4100, 300, 33), (5000, 700, 33), (8192, 64, 33)]


@pytest.mark.parametrize("ql,sl,ns", LENGTHS)
def test_all_geometries_vs_oracle(B, ql, sl, ns):
    rng = np.random.default_rng(ql * 131 + sl)
    q = R.random_rows(rng, 2, ql, with_n=0.01)
    s = R.random_rows(rng, ns, sl, with_n=0.01)
    s[: ns // 3, : min(ql, sl)] = q[0, : min(ql, sl)]
    for algo in (B.MYERS_GLOBAL, B.MYERS_SEMIGLOBAL, B.BITPAL_PACKED):
        assert (gpu(B, algo, q, s) == expect(algo, q, s)).all(), (algo, ql, sl)
    if ql <= 5120:
        assert (gpu(B, B.BITPAL_NONPACKED, q, s) == expect(3, q, s)).all()
    for M, I, G in ((1, -1, -1), (1, -3, -2)):
        kw = dict(match=M, mismatch=I, gap=G)
        assert (gpu(B, B.BITPAL_PACKED, q, s, **kw) == expect(3, q, s, **kw)).all(), (M, I, G, ql, sl)
    if ql <= 2048:
        kw = dict(match=1, mismatch=-3, gap=-2)
        assert (gpu(B, B.BITPAL_NONPACKED, q, s, **kw) == expect(3, q, s, **kw)).all()


@pytest.mark.parametrize("ql,sl,ns", [(150, 150, 500), (100, 400, 333), (37, 150, 65), (150, 37, 65), (500, 800, 100), (1, 9, 33),
                                      (320, 321, 70), (1500, 2500, 40), (5000, 5300, 33), (5000, 700, 33), (64, 64, 64)])
def test_bitpal_semiglobal_vs_oracle_and_dp(B, ql, sl, ns):
    """BitPAl semi-global (whole query inside the subject): no reference build exists
    (the generator needs a JRE) -- parity is against the restated emission (oracle) and, on a sample, plain DP."""
    rng = np.random.default_rng(ql * 17 + sl)
    q = R.random_rows(rng, 2, ql, with_n=0.01)
    s = R.random_rows(rng, ns, sl, with_n=0.01)
    if sl >= ql:
        for i in range(ns // 3):
            off = int(rng.integers(0, sl - ql + 1))
            s[i, off:off + ql] = q[i % 2, :ql]
    for M, I, G in ((2, -3, -5), (1, -1, -1), (1, -3, -2)):
        p = B.Params.default(B.BITPAL_PACKED_SEMIGLOBAL, match=M, mismatch=I, gap=G)
        got = B.align_batch(p, q, s)
        assert (got == R.oracle_batch(R.ALGO_BITPAL_SEMI, q, s, M=M, I=I, G=G)).all(), (M, I, G, ql, sl)
        if ql * sl <= 250_000:
            assert (got[:, :8] == R.dp_scores("nw_semi", q, s[:8], M=M, I=I, G=G)).all()
        if sl >= ql:
            assert (got[0, 0] == M * ql) or ns < 3          # a planted exact copy scores M per base


def test_myers_long_queries(B):
    rng = np.random.default_rng(3)
    for ql, sl in ((16384, 300), (20000, 150), (32768, 100)):
        q = R.random_rows(rng, 1, ql); s = R.random_rows(rng, 33, sl)
        s[0, :sl] = q[0, 5000:5000 + sl]
        for algo in (B.MYERS_GLOBAL, B.MYERS_SEMIGLOBAL):
            assert (gpu(B, algo, q, s) == expect(algo, q, s)).all(), (algo, ql, sl)


@pytest.mark.parametrize("L,e", [(100, 5), (100, 15), (100, 16), (100, 31), (64, 3), (50, 5), (33, 1), (250, 7), (640, 20), (333, 31), (1000, 10)])
def test_banded_vs_oracle(B, L, e):
    rng = np.random.default_rng(L * 37 + e)
    q = R.random_rows(rng, 2, L, with_n=0.005)
    s = np.concatenate([R.mutate_rows(rng, q[0, :L], 300, 2 * e + 2), R.indel_rows(rng, q[1, :L], 200, e + 3),
                        R.random_rows(rng, 77, L, with_n=0.01)])
    got = gpu(B, B.BANDED_MYERS, q, s, threshold=e)
    exp = expect(2, q, s, threshold=e)
    assert (got == exp).all()
    if 4 * e < L:
        assert (exp == 127).any() and (exp < 127).any()


@pytest.mark.parametrize("L,e,policy", [(100, 5, None), (100, 5, "off"), (100, 5, "0,32"), (100, 5, "0,8"), (100, 5, "1,24"),
                                        (150, 5, "2,32"), (250, 7, None), (250, 15, "1,32"), (100, 20, "0,32"), (333, 31, "3,30"),
                                        (640, 20, None), (96, 5, "1,32"), (97, 5, "1,32")])
def test_banded_shuffled_survivor_compaction(B, L, e, policy, monkeypatch):
    """Near-matches and unrelated subjects INTERLEAVED: the kernel parks the survivors of a tile in a shared-memory
    ring after row block P and finishes them 32 at a time (banded.cuh "Survivor compaction"), so the scores -- and,
    bench.py workload C3s, the speed -- must not depend on the order of the subjects.  Every policy (block, max alive;
    off) must give the oracle's scores, through the packed-tile kernel, the fused ASCII kernel and with several
    queries (the ring is flushed when a warp moves on to another query)."""
    import torch
    if policy is None:
        monkeypatch.delenv("BGSA_BANDED_REFILL", raising=False)
    else:
        monkeypatch.setenv("BGSA_BANDED_REFILL", policy)
    rng = np.random.default_rng(L * 101 + e)
    n = 6000 if L <= 150 else 1500
    q = R.random_rows(rng, 3, L)
    s = np.concatenate([R.mutate_rows(rng, q[0, :L], n // 3, 2 * e + 2), R.indel_rows(rng, q[1, :L], n // 6, e + 3),
                        R.mutate_rows(rng, q[2, :L], n // 6, e), R.random_rows(rng, n - n // 3 - 2 * (n // 6), L)])
    s = np.ascontiguousarray(s[rng.permutation(n)])
    s[rng.integers(0, n, 3), rng.integers(0, L, 3)] = ord("N")          # a few tiles take the N path (never parked)
    exp = expect(2, q, s, threshold=e)
    p = B.Params.default(B.BANDED_MYERS, threshold=e)
    for nq in (1, 3):
        got = B.align_batch(p, q[:nq], s)                                # fused kernel (rows <= ~950 bases)
        assert (got == exp[:nq]).all(), ("fused", L, e, policy, nq)
        d_rows = torch.from_numpy(s.reshape(-1)).cuda()
        d_packed = torch.empty(B.packed_bytes(L, n), dtype=torch.uint8, device="cuda")
        d_res = torch.full((nq * n,), 99, dtype=torch.int8, device="cuda")
        B.pack_subjects_device(p, d_rows.data_ptr(), L, n, d_packed.data_ptr())
        B.align_device(p, q[:nq], d_packed.data_ptr(), L, n, d_res.data_ptr(), n)
        torch.cuda.synchronize()
        assert (d_res.cpu().numpy().reshape(nq, n) == exp[:nq]).all(), ("packed", L, e, policy, nq)


# ---- edge cases ------------------------------------------------------------------------------------
def test_empty_and_tiny_batches(B):
    rng = np.random.default_rng(1)
    q = R.random_rows(rng, 1, 150)
    empty = np.zeros((0, 151), np.uint8)
    assert gpu(B, B.BITPAL_PACKED, q, empty).shape == (1, 0)
    for ns in (1, 31, 32, 33):
        s = R.random_rows(rng, ns, 150)
        assert (gpu(B, B.BITPAL_PACKED, q, s) == expect(3, q, s)).all()


def test_subrange_and_result_stride(B):
    rng = np.random.default_rng(2)
    q = R.random_rows(rng, 3, 150); s = R.random_rows(rng, 500, 150)
    p = B.Params.default(B.MYERS_GLOBAL)
    out = np.full((3, 640), 7, np.int16)
    B.align_batch(p, q, s, first=100, count=333, out=out[:, 64:64 + 333])
    exp = expect(0, q, s[100:433])
    assert (out[:, 64:64 + 333] == exp).all() and (out[:, :64] == 7).all() and (out[:, 397:] == 7).all()


def test_non_acgtn_bytes_behave_as_A(B):
    # Appendix A3: mapping_table is zero-initialised, so any other byte is 'A'
    rng = np.random.default_rng(4)
    q = R.random_rows(rng, 1, 100); s = R.random_rows(rng, 64, 100)
    s2 = s.copy(); s2[s2 == ord("A")] = ord("x"); s2[5, 7] = ord("a"); s2[6, 8] = ord("-")
    s_ref = s.copy(); s_ref[5, 7] = ord("A"); s_ref[6, 8] = ord("A")
    for algo in (B.MYERS_GLOBAL, B.BITPAL_PACKED):
        assert (gpu(B, algo, q, s2) == gpu(B, algo, q, s_ref)).all()
    assert (gpu(B, B.BANDED_MYERS, q, s2, threshold=9) == gpu(B, B.BANDED_MYERS, q, s_ref, threshold=9)).all()


def test_int16_wrap_like_reference(B):
    L = 6000
    q = np.full((1, L + 1), ord("A"), np.uint8); q[:, L] = 10
    s = np.full((40, L + 1), ord("C"), np.uint8); s[:, L] = 10
    s[1, :L] = ord("A")
    got = gpu(B, B.MYERS_GLOBAL, q, s)
    assert got[0, 0] == -6000 and got[0, 1] == 0
    # (2,-3,-5) on 6 kbp of pure mismatches: -18000 fits; the narrowing itself is covered by 8 kbp of gaps below
    bp = gpu(B, B.BITPAL_PACKED, q, s)
    assert bp[0, 0] == -18000 and bp[0, 1] == 12000
    q2 = np.full((1, 8001), ord("A"), np.uint8); q2[:, 8000] = 10
    s3 = np.full((33, 11), ord("C"), np.uint8); s3[:, 10] = 10
    bp2 = gpu(B, B.BITPAL_PACKED, q2, s3)          # true score -5*7990 - 3*10 = -39980 -> wraps
    assert bp2[0, 0] == np.int16(np.int32(-39980).astype(np.int16)) and (bp2 == expect(3, q2, s3)).all()


def test_many_queries_layout(B):
    # [query][subject] row-major (cal_cpu.c:79-82), more queries than one ref bucket (REF_BUCKET_COUNT 100)
    rng = np.random.default_rng(5)
    q = R.random_rows(rng, 130, 90); s = R.random_rows(rng, 70, 100)
    for algo in (B.MYERS_GLOBAL, B.BITPAL_PACKED):
        assert (gpu(B, algo, q, s) == expect(algo, q, s)).all()


# ---- full-size runs: size-independent properties + oracle on a sample ------------------------------
def _sample_check(B, algo, q, s, got, rng, n=512, **kw):
    idx = np.sort(rng.choice(s.shape[0], size=n, replace=False))
    assert (got[:, idx] == expect(algo, q, np.ascontiguousarray(s[idx]), **kw)).all()


def test_c2_full_size_properties(B):
    q, s = synth.make("C2")
    assert s.shape == (1_000_000, 151)
    got = gpu(B, B.BITPAL_PACKED, q, s)
    rng = np.random.default_rng(0)
    _sample_check(B, 3, q, s, got, rng)
    # permutation equivariance: scoring a shuffled database gives the shuffled scores
    perm = rng.permutation(s.shape[0])
    assert (gpu(B, B.BITPAL_PACKED, q, np.ascontiguousarray(s[perm])) == got[:, perm]).all()
    # symmetry of the global score: swap the roles of query and subject for a few pairs
    for i in (0, 1, 999_999):
        qi = np.ascontiguousarray(s[i:i + 1]); si = np.ascontiguousarray(q)
        assert gpu(B, B.BITPAL_PACKED, qi, si)[0, 0] == got[0, i]
    # self alignment = match * length; Myers distance of the query to itself = 0
    assert gpu(B, B.BITPAL_PACKED, q, q)[0, 0] == 2 * 150
    assert gpu(B, B.MYERS_GLOBAL, q, q)[0, 0] == 0
    # a checksum of the whole score vector recomputed from two halves (sub-range API)
    p = B.Params.default(B.BITPAL_PACKED)
    a = B.align_batch(p, q, s, first=0, count=400_007)
    b = B.align_batch(p, q, s, first=400_007, count=599_993)
    assert int(a.astype(np.int64).sum() + b.astype(np.int64).sum()) == int(got.astype(np.int64).sum())


def test_c3_full_size_properties(B):
    q, s = synth.make("C3", 2_000_000)
    got = gpu(B, B.BANDED_MYERS, q, s, threshold=5)
    rng = np.random.default_rng(1)
    _sample_check(B, 2, q, s, got, rng, n=2048, threshold=5)
    frac_ok = (got[0, : s.shape[0] // 2] < 127).mean()
    assert 0.5 < frac_ok < 1.0                       # the similar half mostly verifies ...
    assert (got[0, s.shape[0] // 2:] == 127).mean() > 0.999    # ... the random half is rejected
    assert got.min() >= 0


def test_c4_slice_properties(B):
    q, s = synth.make("C4", 20_000)
    got = gpu(B, B.MYERS_SEMIGLOBAL, q, s)
    glob = gpu(B, B.MYERS_GLOBAL, q, s)
    rng = np.random.default_rng(2)
    _sample_check(B, 1, q, s, got, rng, n=96)
    assert (got >= glob).all()                       # -distance: a free query substring can only help
    assert gpu(B, B.MYERS_SEMIGLOBAL, q, np.ascontiguousarray(q))[0, 0] == 0


def test_c5_slice(B):
    q, s = synth.make("C5", 300)
    got = gpu(B, B.BITPAL_PACKED, q, s)
    rng = np.random.default_rng(3)
    _sample_check(B, 3, q, s, got, rng, n=24)
    assert gpu(B, B.BITPAL_PACKED, q, q)[0, 0] == 10000


def test_device_resident_api_matches_host_api(B):
    import torch
    q, s = synth.make("C2", 50_000)
    p = B.Params.default(B.BITPAL_PACKED)
    d_rows = torch.from_numpy(s.reshape(-1)).cuda()
    d_packed = torch.empty(B.packed_bytes(150, s.shape[0]), dtype=torch.uint8, device="cuda")
    d_res = torch.empty(s.shape[0], dtype=torch.int16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    B.pack_subjects_device(p, d_rows.data_ptr(), 150, s.shape[0], d_packed.data_ptr(), 0, st)
    B.align_device(p, q, d_packed.data_ptr(), 150, s.shape[0], d_res.data_ptr(), s.shape[0], 0, st)
    torch.cuda.synchronize()
    assert (d_res.cpu().numpy()[None, :] == B.align_batch(p, q, s)).all()


# ---- pack kernels: the packed tile layout itself (csrc/bgsa_common.cuh), against a numpy packer -----
_numpy_pack = R.numpy_pack


@pytest.mark.parametrize("slen,n", [(150, 1000), (100, 4099), (1000, 130), (15, 77), (16, 64), (31, 33), (127, 97), (511, 65),
                                    (63, 32), (64, 31), (65, 1), (5000, 40), (3, 50), (255, 200), (4095, 37)])
@pytest.mark.parametrize("layout_algo", [0, 2])
def test_pack_kernels_bit_exact(B, slen, n, layout_algo):
    """Both pack kernels (streaming, and the simple one used for tiny / huge rows) against numpy: clean
    rows, rows with N / lower case / arbitrary bytes, and a row buffer that is not 16-byte aligned."""
    import torch
    rng = np.random.default_rng(slen * 7 + n)
    p = B.Params.default(layout_algo, threshold=5)
    layout = 1 if layout_algo == 2 else 0
    for variant in ("clean", "dirty"):
        rows = R.random_rows(rng, n, slen, with_n=0.0 if variant == "clean" else 0.02)
        if variant == "dirty":
            junk = rng.random(rows[:, :slen].shape) < 0.01
            rows[:, :slen][junk] = rng.integers(0, 256, size=int(junk.sum()), dtype=np.uint8)
            rows[::7, slen] = 13          # a row end that is not a newline
            rows[n // 2, :slen] = ord("A")
        for shift in (0, 5):
            buf = torch.zeros(rows.size + 64, dtype=torch.uint8, device="cuda")
            buf[shift:shift + rows.size] = torch.from_numpy(rows.reshape(-1)).cuda()
            nbytes = B.packed_bytes(slen, n)
            d_packed = torch.full((nbytes,), 0xA5, dtype=torch.uint8, device="cuda")
            B.pack_subjects_device(p, buf.data_ptr() + shift, slen, n, d_packed.data_ptr())
            torch.cuda.synchronize()
            raw = d_packed.cpu().numpy()
            ntiles, ku, kn = (n + 31) // 32, (slen + 63) // 64, (slen + 31) // 32
            up = lambda x: (x + 255) // 256 * 256
            codes = raw[: ntiles * ku * 512].view(np.uint32).reshape(ntiles, ku, 32, 4)
            o1 = up(ntiles * ku * 512)
            nm = raw[o1: o1 + ntiles * kn * 128].view(np.uint32).reshape(ntiles, kn, 32)
            flags = raw[o1 + up(ntiles * kn * 128): o1 + up(ntiles * kn * 128) + ntiles]
            ec, en, ef = _numpy_pack(rows, layout)
            assert (codes == ec).all(), (variant, shift)
            assert ((flags != 0) == ef).all(), (variant, shift)
            assert (nm[ef] == en[ef]).all(), (variant, shift)      # the N plane is only defined for flagged tiles


def test_host_register_pins_caller_buffers(B):
    """bgsa_host_register: the integrator keeps its own (malloc'ed) buffers; results are unchanged."""
    import ctypes as C
    q, s = synth.make("C2", 70_000)
    lib = B.load()
    s2 = np.ascontiguousarray(s.copy())
    out = np.zeros((1, s2.shape[0]), dtype=np.int16)
    assert lib.bgsa_host_register(s2.ctypes.data, s2.nbytes) == 0
    assert lib.bgsa_host_register(out.ctypes.data, out.nbytes) == 0
    try:
        got = B.align_batch(B.Params.default(B.BITPAL_PACKED), q, s2, out=out)
    finally:
        assert lib.bgsa_host_unregister(s2.ctypes.data) == 0
        assert lib.bgsa_host_unregister(out.ctypes.data) == 0
    assert (got == expect(3, q, s)).all()
    assert lib.bgsa_host_register(None, 10) == 1          # BGSA_ERR_ARG


def test_seeded_fuzz_all_algorithms(B):
    """80 seeded random (query length, subject length, count, N rate) cases per algorithm family against the
    oracle: odd lengths, counts off the tile/pass/chunk grids, lengths around every instance boundary."""
    rng = np.random.default_rng(2026)
    edges = [31, 32, 33, 63, 64, 65, 95, 96, 97, 159, 160, 161, 191, 192, 193, 255, 256, 257, 319, 320, 321, 383, 385, 639, 641]
    for case in range(80):
        ql = int(rng.choice(edges)) if case % 3 == 0 else int(rng.integers(1, 700))
        sl = int(rng.choice(edges)) if case % 5 == 0 else int(rng.integers(1, 700))
        ns = int(rng.integers(1, 400))
        q = R.random_rows(rng, int(rng.integers(1, 4)), ql, with_n=float(rng.choice([0.0, 0.01, 0.2])))
        s = R.random_rows(rng, ns, sl, with_n=float(rng.choice([0.0, 0.0, 0.02])))
        k = min(ql, sl)
        s[: ns // 2, :k] = q[0, :k]
        for algo, oalgo in ((B.MYERS_GLOBAL, 0), (B.MYERS_SEMIGLOBAL, 1), (B.BITPAL_PACKED, 3), (B.BITPAL_PACKED_SEMIGLOBAL, 5)):
            got = B.align_batch(B.Params.default(algo), q, s)
            assert (got == R.oracle_batch(oalgo, q, s)).all(), (case, algo, ql, sl, ns)
        if case % 4 == 0:
            assert (B.align_batch(B.Params.default(B.BITPAL_NONPACKED), q, s) == R.oracle_batch(3, q, s)).all(), (case, ql, sl, ns)
        e = int(rng.integers(1, 16))
        if sl >= 20 and ((sl - 1) // 64 + 1) < ((sl - e + 63) // 64 + 1):     # banded: equal lengths, reference in bounds
            qb = R.random_rows(rng, 1, sl)
            sb = np.concatenate([R.mutate_rows(rng, qb[0, :sl], ns, 2 * e), R.random_rows(rng, 7, sl)])
            got = B.align_batch(B.Params.default(B.BANDED_MYERS, threshold=e), qb, sb)
            assert (got == R.oracle_batch(R.ALGO_BANDED, qb, sb, e=e)).all(), (case, "banded", sl, e)


def test_two_slots_in_flight(B):
    """bgsa_align_batch_submit/_wait: two jobs (slot 0 and 1, the reference's a/b ping-pong buffers, thread.c:35-170)
    queued back to back on one device with different algorithms and pinned buffers from bgsa_malloc_host."""
    import ctypes as C
    lib = B.load()
    rng = np.random.default_rng(12)
    qa = R.random_rows(rng, 2, 150); sa = R.random_rows(rng, 50_000, 150)
    qb = R.random_rows(rng, 1, 300); sb = R.random_rows(rng, 20_011, 260, with_n=0.01)

    def pinned_copy(arr):
        ptr = lib.bgsa_malloc_host(arr.nbytes)
        assert ptr
        buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(arr.nbytes,))
        buf[:] = arr.reshape(-1).view(np.uint8)
        return ptr, buf

    pa, ba = pinned_copy(sa); pb, bb = pinned_copy(sb)
    ra_ptr = lib.bgsa_malloc_host(2 * 2 * sa.shape[0]); rb_ptr = lib.bgsa_malloc_host(2 * sb.shape[0])
    try:
        seq_a = B.SeqT(150, sa.nbytes, sa.shape[0], 0, 0, pa); seq_b = B.SeqT(260, sb.nbytes, sb.shape[0], 0, 0, pb)
        qca = np.ascontiguousarray(B.to_codes(qa)); qcb = np.ascontiguousarray(B.to_codes(qb))
        p_a = B.Params.default(B.BITPAL_PACKED); p_b = B.Params.default(B.MYERS_SEMIGLOBAL)
        assert lib.bgsa_align_batch_submit(C.byref(p_a), qca.ctypes.data, 2, 150, C.byref(seq_a), 0, sa.shape[0], ra_ptr, sa.shape[0], 0, 0) == 0
        assert lib.bgsa_align_batch_submit(C.byref(p_b), qcb.ctypes.data, 1, 300, C.byref(seq_b), 11, sb.shape[0] - 11, rb_ptr, sb.shape[0] - 11, 0, 1) == 0
        assert lib.bgsa_align_batch_wait(0, 1) == 0
        assert lib.bgsa_align_batch_wait(0, 0) == 0
        ra = np.ctypeslib.as_array(C.cast(ra_ptr, C.POINTER(C.c_int16)), shape=(2, sa.shape[0])).copy()
        rb = np.ctypeslib.as_array(C.cast(rb_ptr, C.POINTER(C.c_int16)), shape=(1, sb.shape[0] - 11)).copy()
        assert lib.bgsa_align_batch_wait(0, 2) == 1 and lib.bgsa_align_batch_submit(C.byref(p_a), qca.ctypes.data, 2, 150, C.byref(seq_a), 0, 1, ra_ptr, 1, 0, 2) == 1
    finally:
        for p in (pa, pb, ra_ptr, rb_ptr):
            lib.bgsa_free_host(p)
    assert (ra == R.oracle_batch(R.ALGO_BITPAL_PACKED, qa, sa)).all()
    assert (rb == R.oracle_batch(R.ALGO_MYERS_SEMIGLOBAL, qb, sb[11:])).all()


@pytest.mark.parametrize("nq,ns", [(300, 70), (7, 5000), (101, 1), (1000, 33)])
def test_query_counts_beyond_the_grid(B, nq, ns):
    """More queries than CTAs, queries nobody starts on, one subject: the CTAs' walk over the queries must
    visit every (query, tile) exactly once."""
    rng = np.random.default_rng(nq * 31 + ns)
    q = R.random_rows(rng, nq, 90, with_n=0.01)
    s = R.random_rows(rng, ns, 75, with_n=0.01)
    for algo, oalgo in ((B.MYERS_GLOBAL, 0), (B.BITPAL_PACKED, 3), (B.MYERS_SEMIGLOBAL, 1)):
        assert (B.align_batch(B.Params.default(algo), q, s) == R.oracle_batch(oalgo, q, s)).all(), (algo, nq, ns)
    qb = R.random_rows(rng, nq, 80)
    sb = np.concatenate([R.mutate_rows(rng, qb[0, :80], ns, 10), R.random_rows(rng, 5, 80)])
    got = B.align_batch(B.Params.default(B.BANDED_MYERS, threshold=5), qb, sb)
    assert (got == R.oracle_batch(R.ALGO_BANDED, qb, sb, e=5)).all()
    if nq == 300:    # a wavefront instance as well (L > 1)
        ql = R.random_rows(rng, 40, 700); sl = R.random_rows(rng, 50, 200)
        assert (B.align_batch(B.Params.default(B.BITPAL_PACKED), ql, sl) == R.oracle_batch(3, ql, sl)).all()


@pytest.mark.parametrize("slen,n,e", [(100, 5000, 5), (100, 4099, 15), (64, 333, 3), (50, 97, 5), (250, 1000, 7), (333, 100, 31),
                                      (640, 200, 20), (900, 70, 10), (1000, 60, 10), (33, 64, 1), (15, 40, 2)])
def test_rows_device_entry_banded_fused(B, slen, n, e):
    """bgsa_align_rows_device: ASCII rows resident in HBM -> scores.  For banded Myers this is the fused kernel (tile
    encoded into shared memory, band run from there); rows with N / arbitrary bytes, several queries, a row buffer
    that is not 16-byte aligned, counts off the tile grid.  (slen 1000: falls back to pack + align.)"""
    import torch
    rng = np.random.default_rng(slen * 13 + n)
    q = R.random_rows(rng, 3, slen, with_n=0.004)
    s = np.concatenate([R.mutate_rows(rng, q[0, :slen], n // 2, 2 * e + 2), R.indel_rows(rng, q[1, :slen], n // 4, e + 3),
                        R.random_rows(rng, n - n // 2 - n // 4, slen, with_n=0.01)])
    junk = rng.random(s[:, :slen].shape) < 0.002
    s[:, :slen][junk] = rng.integers(0, 256, size=int(junk.sum()), dtype=np.uint8)
    p = B.Params.default(B.BANDED_MYERS, threshold=e)
    exp = R.oracle_batch(R.ALGO_BANDED, q, s, e=e)
    inb = ((slen - 1) // 64 + 1) < ((slen - e + 63) // 64 + 1)
    for shift in (0, 7):
        buf = torch.zeros(s.size + 64, dtype=torch.uint8, device="cuda")
        buf[shift:shift + s.size] = torch.from_numpy(s.reshape(-1)).cuda()
        d_res = torch.full((3 * n,), 99, dtype=torch.int8, device="cuda")
        before = B.launch_count()
        B.align_rows_device(p, q, buf.data_ptr() + shift, slen, n, d_res.data_ptr(), n)
        torch.cuda.synchronize()
        got = d_res.cpu().numpy().reshape(3, n)
        assert B.launch_count() - before == (1 if slen <= 900 else 2)        # fused: one kernel
        assert (got == B.align_batch(p, q, s)).all()                         # the two-kernel path agrees
        if inb:
            assert (got == exp).all(), (slen, n, e, shift)


def test_rows_device_entry_other_algorithms(B):
    import torch
    rng = np.random.default_rng(77)
    q = R.random_rows(rng, 2, 150, with_n=0.01); s = R.random_rows(rng, 3001, 170, with_n=0.01)
    d_rows = torch.from_numpy(s.reshape(-1)).cuda()
    for algo, oalgo in ((B.MYERS_GLOBAL, 0), (B.MYERS_SEMIGLOBAL, 1), (B.BITPAL_PACKED, 3), (B.BITPAL_PACKED_SEMIGLOBAL, 5)):
        d_res = torch.zeros(2 * 3001 * 2, dtype=torch.uint8, device="cuda")
        B.align_rows_device(B.Params.default(algo), q, d_rows.data_ptr(), 170, 3001, d_res.data_ptr(), 3001)
        torch.cuda.synchronize()
        assert (d_res.cpu().numpy().view(np.int16).reshape(2, 3001) == R.oracle_batch(oalgo, q, s)).all(), algo


@pytest.mark.parametrize("ql,sl,n", [(150, 150, 4099), (100, 100, 1000), (64, 37, 333), (256, 250, 700), (33, 400, 97), (1, 9, 70),
                                     (150, 151, 64), (200, 135, 31), (96, 15, 65), (250, 303, 2049), (330, 300, 200), (384, 250, 97)])
def test_rows_kernel_every_algorithm(B, ql, sl, n, monkeypatch):
    """align_rows_kernel (rows_kernel.cuh): thread per subject straight from the ASCII rows, match masks looked up by byte
    value.  Against the oracle and against the pack + align path, for every algorithm that has it; rows with N, lower
    case and arbitrary bytes (global.c:9-15: anything else is an A), a row buffer that is not 16-byte aligned and ends
    exactly at the end of its allocation's last row (nothing may be read past it), counts off the tile grid, one and
    two stages per warp."""
    import torch
    rng = np.random.default_rng(ql * 977 + sl * 13 + n)
    q = R.random_rows(rng, 2, ql, with_n=0.01)
    s = R.random_rows(rng, n, sl, with_n=0.01)
    m = min(ql, sl)
    s[: n // 3, :m] = q[0, :m]
    junk = rng.random(s[:, :sl].shape) < 0.004
    s[:, :sl][junk] = rng.integers(0, 256, size=int(junk.sum()), dtype=np.uint8)
    s[::5, sl] = 0                                            # row ends need not be newlines
    clean = s.copy()
    body = clean[:, :sl]
    body[~np.isin(body, np.frombuffer(b"ACGTN", dtype=np.uint8))] = ord("A")
    algos = [(B.MYERS_GLOBAL, 0, {}), (B.MYERS_SEMIGLOBAL, 1, {}), (B.BITPAL_PACKED, 3, {}), (B.BITPAL_NONPACKED, 3, {}),
             (B.BITPAL_PACKED_SEMIGLOBAL, 5, {}), (B.BITPAL_PACKED, 3, dict(match=1, mismatch=-1, gap=-1)),
             (B.BITPAL_NONPACKED, 3, dict(match=1, mismatch=-3, gap=-2))]
    for algo, oalgo, kw in algos:
        p = B.Params.default(algo, **kw)
        name, fused = B.rows_kernel_name(p, ql, sl)
        if ",L=1>" not in B.kernel_name(p, ql, sl):           # no thread-per-subject instance this wide (non-packed above K = 5,
            assert not fused                                  # packed above K = 10): a wavefront instance on packed tiles serves it
            assert (B.align_batch(p, q, s) == R.oracle_batch(oalgo, q, clean, M=kw.get("match", 2), I=kw.get("mismatch", -3), G=kw.get("gap", -5))).all()
            continue
        assert fused and name.startswith("align_rows_kernel<"), name
        exp = R.oracle_batch(oalgo, q, clean, M=kw.get("match", 2), I=kw.get("mismatch", -3), G=kw.get("gap", -5))
        for shift, stages in ((0, None), (5, "1"), (9, "2")):
            if stages is None:
                monkeypatch.delenv("BGSA_ROWS_STAGES", raising=False)
            else:
                monkeypatch.setenv("BGSA_ROWS_STAGES", stages)
            buf = torch.zeros(shift + s.size, dtype=torch.uint8, device="cuda")   # the rows end where the allocation's data ends
            buf[shift:] = torch.from_numpy(s.reshape(-1)).cuda()
            d_res = torch.full((2 * n,), 12345, dtype=torch.int16, device="cuda")
            before = B.launch_count()
            B.align_rows_device(p, q, buf.data_ptr() + shift, sl, n, d_res.data_ptr(), n)
            torch.cuda.synchronize()
            assert B.launch_count() - before == 1                              # one kernel, no pack launch
            got = d_res.cpu().numpy().reshape(2, n)
            assert (got == exp).all(), (algo, kw, ql, sl, n, shift, stages)
        monkeypatch.delenv("BGSA_ROWS_STAGES", raising=False)
        monkeypatch.setenv("BGSA_NO_ROWS_KERNEL", "1")                         # the same call through pack + align
        assert not B.rows_kernel_name(p, ql, sl)[1]
        assert (B.align_batch(p, q, s) == exp).all(), (algo, kw, "packed path")
        monkeypatch.delenv("BGSA_NO_ROWS_KERNEL")
        assert (B.align_batch(p, q, s) == exp).all(), (algo, kw, "batch entry, rows kernel")


def test_rows_kernel_eligibility(B):
    """Row pitches whose lanes would pile up on a few shared-memory banks, long rows and long queries take pack + align."""
    p = B.Params.default(B.MYERS_GLOBAL)
    assert B.rows_kernel_name(p, 150, 150)[1]
    assert not B.rows_kernel_name(p, 150, 127)[1]        # pitch 128: every lane on the same bank
    assert not B.rows_kernel_name(p, 150, 63)[1]
    assert not B.rows_kernel_name(p, 150, 1000)[1]       # a tile does not fit the stage
    assert B.rows_kernel_name(p, 384, 150)[1]            # 12 words: the largest rows instance
    assert not B.rows_kernel_name(p, 500, 150)[1]        # K > 12: no rows instance
    rng = np.random.default_rng(3)
    q = R.random_rows(rng, 1, 150); s = R.random_rows(rng, 500, 127, with_n=0.01)
    assert (B.align_batch(p, q, s) == R.oracle_batch(0, q, s)).all()


JIT_SCHEMES = [(0, -1, -1), (4, -6, -10), (5, -3, -4), (1, -1, -2), (3, -2, -4), (2, -1, -1), (1, -2, -3), (6, -1, -7), (10, -15, -25)]


@pytest.mark.parametrize("M,I,G", JIT_SCHEMES)
def test_unlisted_scoring_schemes_instantiated_at_run_time(B, M, I, G, tmp_path_factory, monkeypatch):
    """Any valid (match, mismatch, gap) runs: schemes outside the compiled list are instantiated by NVRTC from the embedded
    kernel headers (csrc/jit.cu) -- what the reference does by re-running its generator (Main.java:240-315; the (0,-1,-1)
    "edit" special case :270-272; the common-factor reduction :213-267, here (4,-6,-10) and (10,-15,-25)).  Packed,
    non-packed and semi-global, thread-per-subject (rows kernel and packed tiles) and wavefront geometries, against plain
    DP and the restated BitPAl (oracle)."""
    monkeypatch.setenv("BGSA_JIT_CACHE", str(tmp_path_factory.getbasetemp() / "jit"))     # shared by the parametrised cases
    rng = np.random.default_rng(1000 + M * 100 - I * 10 - G)
    # (query, subject, count, variants): every instance is one NVRTC compile (1-10 s), so the full sweep of geometries runs
    # for two schemes and the others take the thread-per-subject kernels plus one wavefront instance
    P, N, S = B.BITPAL_PACKED, B.BITPAL_NONPACKED, B.BITPAL_PACKED_SEMIGLOBAL
    if (M, I, G) in JIT_SCHEMES[:2]:
        geos = [(150, 150, 200, (P, N, S)), (60, 90, 70, (P, N, S)), (300, 127, 64, (P,)), (700, 300, 40, (P, S)), (2100, 150, 33, (P,))]
    else:
        geos = [(150, 150, 200, (P, N, S)), (700, 300, 40, (P,))]
    for ql, sl, ns, variants in geos:
        q = R.random_rows(rng, 2, ql, with_n=0.01)
        s = R.random_rows(rng, ns, sl, with_n=0.01)
        m = min(ql, sl)
        s[: ns // 3, :m] = q[0, :m]
        s[ns // 3: ns // 2, :m] = R.mutate_rows(rng, q[1, :m], ns // 2 - ns // 3, max(1, m // 12))[:, :m]
        kw = dict(match=M, mismatch=I, gap=G)
        dp = R.dp_scores("nw", q, s[:6], M=M, I=I, G=G).astype(np.int16)
        for algo in (P, N):
            if algo not in variants:
                continue
            p = B.Params.default(algo, **kw)
            assert B.kernel_name(p, ql, sl).endswith("[NVRTC]")
            got = B.align_batch(p, q, s)
            assert (got == R.oracle_batch(R.ALGO_BITPAL_PACKED, q, s, M=M, I=I, G=G)).all(), (algo, M, I, G, ql, sl)
            assert (got[:, :6] == dp).all(), (algo, M, I, G, ql, sl)
        if S in variants:
            p = B.Params.default(S, **kw)
            got = B.align_batch(p, q, s)
            assert (got == R.oracle_batch(R.ALGO_BITPAL_SEMI, q, s, M=M, I=I, G=G)).all(), ("semi", M, I, G, ql, sl)
            assert (got[:, :6] == R.dp_scores("nw_semi", q, s[:6], M=M, I=I, G=G).astype(np.int16)).all()
    # the device-resident entries take the same route (packed tiles, and the rows kernel on its own)
    import torch
    q = R.random_rows(rng, 1, 150); s = R.random_rows(rng, 4097, 150, with_n=0.01)
    p = B.Params.default(B.BITPAL_PACKED, match=M, mismatch=I, gap=G)
    d_rows = torch.from_numpy(s.reshape(-1)).cuda()
    d_res = torch.zeros(4097, dtype=torch.int16, device="cuda")
    B.align_rows_device(p, q, d_rows.data_ptr(), 150, 4097, d_res.data_ptr(), 4097)
    torch.cuda.synchronize()
    exp = R.oracle_batch(R.ALGO_BITPAL_PACKED, q, s, M=M, I=I, G=G)
    assert (d_res.cpu().numpy().reshape(1, -1) == exp).all()
    d_packed = torch.empty(B.packed_bytes(150, 4097), dtype=torch.uint8, device="cuda")
    d_res.zero_()
    B.pack_subjects_device(p, d_rows.data_ptr(), 150, 4097, d_packed.data_ptr())
    B.align_device(p, q, d_packed.data_ptr(), 150, 4097, d_res.data_ptr(), 4097)
    torch.cuda.synchronize()
    assert (d_res.cpu().numpy().reshape(1, -1) == exp).all()


@pytest.mark.parametrize("force", ["1", "0"])
def test_batch_entry_host_pack_front_end(B, force, monkeypatch):
    """bgsa_align_batch with the subjects encoded by the host threads (BGSA_HOST_PACK=1: csrc/host_pack.cpp, a quarter of
    the bytes over PCIe) and by the device pack kernel (=0) give the same scores as the oracle: every algorithm, rows with
    N / arbitrary bytes, pageable (unpinned) subject memory, a sub-range that does not start on a tile, several queries,
    more subjects than one chunk."""
    monkeypatch.setenv("BGSA_HOST_PACK", force)
    rng = np.random.default_rng(4242)
    q = R.random_rows(rng, 2, 150, with_n=0.01)
    s = R.random_rows(rng, 70_001, 150, with_n=0.001)
    junk = rng.random(s[:, :150].shape) < 0.0005
    s[:, :150][junk] = rng.integers(0, 256, size=int(junk.sum()), dtype=np.uint8)
    s[5, :150] = q[0, :150]
    for algo, oalgo in ((B.MYERS_GLOBAL, 0), (B.MYERS_SEMIGLOBAL, 1), (B.BITPAL_PACKED, 3), (B.BITPAL_NONPACKED, 3), (B.BITPAL_PACKED_SEMIGLOBAL, 5)):
        p = B.Params.default(algo)
        exp = R.oracle_batch(oalgo, q, s)
        assert (B.align_batch(p, q, s) == exp).all(), (algo, force)
        assert (B.align_batch(p, q, s, first=37, count=50_003) == exp[:, 37:37 + 50_003]).all(), (algo, force)
    qb = R.random_rows(rng, 2, 100)
    sb = np.concatenate([R.mutate_rows(rng, qb[0, :100], 40_000, 10), R.random_rows(rng, 30_001, 100, with_n=0.001)])
    sb = np.ascontiguousarray(sb[rng.permutation(sb.shape[0])])
    for e in (5, 20):
        pb = B.Params.default(B.BANDED_MYERS, threshold=e)
        expb = R.oracle_batch(R.ALGO_BANDED, qb, sb, e=e)
        assert (B.align_batch(pb, qb, sb) == expb).all(), (e, force)
        assert (B.align_batch(pb, qb, sb, first=33, count=60_000) == expb[:, 33:33 + 60_000]).all(), (e, force)
    ql = R.random_rows(rng, 1, 2000); sl = R.random_rows(rng, 9_000, 1200)          # a wavefront instance, long rows
    assert (B.align_batch(B.Params.default(B.BITPAL_PACKED), ql, sl) == R.oracle_batch(3, ql, sl)).all()


def test_front_end_tuner_never_changes_scores(B, monkeypatch):
    """The batch entry tunes, job by job, which share of the chunks the host threads pack (api.cu front end): whatever the
    search tries -- all ASCII, all host-packed, any mix, trial jobs -- the scores are the same, and the share it reports
    stays within [0, 1].  Also the forced modes and the untuned model."""
    monkeypatch.delenv("BGSA_HOST_PACK", raising=False)
    monkeypatch.setenv("BGSA_HOST_PACK_TUNING", "1")         # (a rank alone on its host keeps the static model: force the search on)
    q, s = synth.make("C2", 60_000)
    for algo, kw, oalgo in ((B.MYERS_GLOBAL, {}, 0), (B.BANDED_MYERS, {"threshold": 5}, 2)):
        if algo == B.BANDED_MYERS:
            q, s = synth.make("C3", 120_000)
        p = B.Params.default(algo, **kw)
        import torch
        h = torch.from_numpy(s.reshape(-1)).pin_memory()
        sp = h.numpy().reshape(s.shape)
        exp = R.oracle_batch(oalgo, q, s[:20000], e=5)
        seen = set()
        for i in range(14):
            got = B.align_batch(p, q, sp)
            share = B.batch_front_end(0, 0)
            assert 0.0 <= share <= 1.0
            seen.add(round(share, 3))
            assert (got[:, :20000] == exp).all(), (algo, i, share)
        if B.host_pack_info()[0] >= 6:                    # (a pool too small to matter is never used: nothing to search)
            assert len(seen) >= 2, seen                   # the search really moved
        for mode in ("0", "1", "2"):
            monkeypatch.setenv("BGSA_HOST_PACK", mode)
            assert (B.align_batch(p, q, sp)[:, :20000] == exp).all(), (algo, mode)
            assert B.batch_front_end(0, 0) == {"0": 0.0, "1": 1.0, "2": -1.0}[mode]
        monkeypatch.delenv("BGSA_HOST_PACK")
        monkeypatch.setenv("BGSA_HOST_PACK_NO_TUNING", "1")
        assert (B.align_batch(p, q, sp)[:, :20000] == exp).all()
        assert B.batch_front_end(0, 0) in (0.0, 1.0, -1.0)   # the model's choice: never / always / chunk by chunk
        monkeypatch.delenv("BGSA_HOST_PACK_NO_TUNING")


def test_resident_entries_concurrent_streams_and_alignment(B):
    """bgsa_align_device from several host threads on several streams at once, each with its OWN queries: the
    query tables, work counters and packed scratch are per caller stream, so nothing is shared.  Misaligned
    device pointers are rejected with BGSA_ERR_ARG instead of faulting the context."""
    import threading
    import torch
    rng = np.random.default_rng(2024)
    n, L = 20000, 150
    s = R.random_rows(rng, n, L)
    p = B.Params.default(B.MYERS_GLOBAL)
    d_rows = torch.from_numpy(s.reshape(-1)).cuda()
    d_packed = torch.empty(B.packed_bytes(L, n), dtype=torch.uint8, device="cuda")
    B.pack_subjects_device(p, d_rows.data_ptr(), L, n, d_packed.data_ptr())
    torch.cuda.synchronize()
    nthreads, reps = 4, 6
    queries = [R.random_rows(rng, 1 + t % 2, L) for t in range(nthreads)]
    want = [R.oracle_batch(R.ALGO_MYERS_GLOBAL, q, s) for q in queries]
    errors = []

    def worker(t):
        try:
            torch.cuda.set_device(0)
            stream = torch.cuda.Stream()
            nq = queries[t].shape[0]
            d_res = torch.zeros(nq * n, dtype=torch.int16, device="cuda")
            for r in range(reps):
                algo_p = p if r % 2 == 0 else B.Params.default(B.BITPAL_PACKED)
                exp = want[t] if r % 2 == 0 else R.oracle_batch(R.ALGO_BITPAL_PACKED, queries[t], s)
                if r % 3 == 2:     # the rows entry too (packs into the stream's own scratch)
                    B.align_rows_device(algo_p, queries[t], d_rows.data_ptr(), L, n, d_res.data_ptr(), n, 0, stream.cuda_stream)
                else:
                    B.align_device(algo_p, queries[t], d_packed.data_ptr(), L, n, d_res.data_ptr(), n, 0, stream.cuda_stream)
                stream.synchronize()
                if not (d_res.cpu().numpy().reshape(nq, n) == exp).all():
                    errors.append((t, r))
        except Exception as exc:      # noqa: BLE001
            errors.append((t, repr(exc)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    # more streams than slots: recycling keeps working
    for i in range(12):
        st = torch.cuda.Stream()
        d_res = torch.zeros(n, dtype=torch.int16, device="cuda")
        B.align_device(p, queries[0][:1], d_packed.data_ptr(), L, n, d_res.data_ptr(), n, 0, st.cuda_stream)
        st.synchronize()
        assert (d_res.cpu().numpy() == want[0][0]).all()
    d_res = torch.zeros(n + 8, dtype=torch.int16, device="cuda")
    with pytest.raises(B.BgsaError) as ei:
        B.align_device(p, queries[0][:1], d_packed.data_ptr() + 16, L, n - 32, d_res.data_ptr(), n)
    assert ei.value.code == 1
    with pytest.raises(B.BgsaError) as ei:
        B.align_device(p, queries[0][:1], d_packed.data_ptr(), L, n, d_res.data_ptr() + 1, n)
    assert ei.value.code == 1
    B.align_device(p, queries[0][:1], d_packed.data_ptr(), L, n, d_res.data_ptr() + 2, n)      # element-aligned is enough
    torch.cuda.synchronize()
    assert (d_res.cpu().numpy()[1:n + 1] == want[0][0]).all()


def test_device_management_entries(B):
    import ctypes as C
    import torch
    lib = B.load()
    n = torch.cuda.device_count()
    assert lib.bgsa_init_devices(n) == 0                       # all contexts, in parallel
    assert lib.bgsa_init_devices(0) == 1 and lib.bgsa_init_devices(n + 1) == 1      # BGSA_ERR_ARG
    assert b"requested" in lib.bgsa_last_error()
    node = C.c_int(-7)
    assert lib.bgsa_bind_thread_to_device(0, C.byref(node)) == 0 and node.value >= -1
    q, s = synth.make("C2", 500)                                # the thread is still usable afterwards
    assert (B.align_batch(B.Params.default(B.MYERS_GLOBAL), q, s) == expect(0, q, s)).all()
