"""CPU tier: pins the oracle (oracle/bgsa_oracle.c) against
  (1) the reference's only checked-in golden output (banded/BGSA_KNC/data/result.txt),
  (2) vectors produced by the unmodified reference builds (tests/golden/ref_vectors.npz, made by
      tests/golden/make_golden.py from oracle/_ref),
  (3) the reference itself, live, when oracle/_ref is present and the CPU supports it,
  (4) plain O(nm) DP.
"""
import ctypes as C
import hashlib

import numpy as np
import pytest

import refutil as R


@pytest.fixture(scope="module")
def vec(golden_dir):
    return np.load(golden_dir / "ref_vectors.npz")


def test_semiglobal_matches_checked_in_golden(golden_dir):
    g = np.load(golden_dir / "golden_semiglobal_knc.npz")
    assert hashlib.md5(g["raw"].tobytes()).hexdigest() == "25ea1bf2962d312b94528e3fde0f0656"
    q, s = R.sample_data()
    got = R.oracle_batch(R.ALGO_MYERS_SEMIGLOBAL, q, s)
    assert got.shape == (3, 128)
    assert (got == g["scores"]).all()
    assert got[0, :5].tolist() == [-249, -256, -244, -254, -257]


def test_sample_data_known_answers(vec):
    q, s = R.sample_data()
    myers = R.oracle_batch(R.ALGO_MYERS_GLOBAL, q, s)
    # result.bin md5 of original/BGSA_CPU on sample-data
    assert hashlib.md5(myers.tobytes()).hexdigest() == "7253c1f2a6423aaa3e29577acc137302"
    assert (myers == vec["sample_myers_cpu"]).all() and (myers == vec["sample_myers_sse"]).all()
    bitpal = R.oracle_batch(R.ALGO_BITPAL_PACKED, q, s)
    assert hashlib.md5(bitpal.tobytes()).hexdigest().startswith("6dcc0519e8")
    assert (bitpal == vec["sample_bitpal_avx512"]).all() and (bitpal == vec["sample_bitpal_avx2"]).all()
    assert (R.oracle_batch(R.ALGO_BITPAL_NONPACKED, q, s) == bitpal).all()
    assert (R.oracle_batch(R.ALGO_BANDED, q, s, e=31) == vec["sample_banded_k31"]).all()


@pytest.mark.parametrize("name,algo,key,kw", [
    ("C2", R.ALGO_BITPAL_PACKED, "C2_ref", {}),
    ("C2", R.ALGO_BITPAL_NONPACKED, "C2_ref", {}),
    ("C2", R.ALGO_MYERS_GLOBAL, "C2_myers_ref", {}),
    ("C3", R.ALGO_BANDED, "C3_ref", {"e": 5}),
    ("C4", R.ALGO_MYERS_SEMIGLOBAL, "C4_ref_restated", {}),
    ("C4", R.ALGO_MYERS_GLOBAL, "C4_myers_ref", {}),
    ("C5", R.ALGO_BITPAL_PACKED, "C5_ref", {}),
])
def test_config_slices_match_reference_vectors(vec, name, algo, key, kw):
    got = R.oracle_batch(algo, vec[f"{name}_query"], vec[f"{name}_subjects"], **kw)
    assert (got == vec[key]).all()


def test_c3_slice_is_discriminating(vec):
    ref = vec["C3_ref"][0]
    assert (ref == 127).sum() > 1000 and (ref < 127).sum() > 1000 and len(np.unique(ref)) >= 8


@pytest.mark.parametrize("i", range(5))
def test_ragged_and_n_cases(vec, i):
    q, s = vec[f"rag{i}_query"], vec[f"rag{i}_subjects"]
    assert (R.oracle_batch(R.ALGO_MYERS_GLOBAL, q, s) == vec[f"rag{i}_myers"]).all()
    assert (R.oracle_batch(R.ALGO_BITPAL_PACKED, q, s) == vec[f"rag{i}_bitpal"]).all()
    assert (R.oracle_batch(R.ALGO_BITPAL_NONPACKED, q, s) == vec[f"rag{i}_bitpal"]).all()


def test_dp_equivalence_random():
    rng = np.random.default_rng(7)
    for _ in range(12):
        ql, sl = int(rng.integers(1, 200)), int(rng.integers(1, 200))
        q = R.random_rows(rng, 2, ql, with_n=0.05)
        s = R.random_rows(rng, 9, sl, with_n=0.05)
        s[:3, : min(ql, sl)] = q[0, : min(ql, sl)]
        assert (R.oracle_batch(R.ALGO_MYERS_GLOBAL, q, s) == -R.dp_scores("edit", q, s)).all()
        assert (R.oracle_batch(R.ALGO_MYERS_SEMIGLOBAL, q, s) == -R.dp_scores("semi", q, s)).all()
        # every scheme the GPU tier runs: the three built in, and those instantiated at run time
        # (tests/test_gpu_parity.py JIT_SCHEMES: the edit special case, common factors 2 and 5, wide and narrow delta ranges)
        for M, I, G in ((2, -3, -5), (1, -1, -1), (1, -3, -2), (4, -6, -10), (5, -4, -3), (0, -1, -1), (5, -3, -4), (1, -1, -2),
                        (3, -2, -4), (2, -1, -1), (1, -2, -3), (6, -1, -7), (10, -15, -25)):
            dp = R.dp_scores("nw", q, s, M, I, G)
            assert (R.oracle_batch(R.ALGO_BITPAL_PACKED, q, s, M=M, I=I, G=G) == dp).all(), (M, I, G)
            assert (R.oracle_batch(R.ALGO_BITPAL_NONPACKED, q, s, M=M, I=I, G=G) == dp).all(), (M, I, G)
            assert (R.oracle_batch(R.ALGO_BITPAL_SEMI, q, s, M=M, I=I, G=G) == R.dp_scores("nw_semi", q, s, M, I, G)).all(), (M, I, G)


def test_int16_wrap_like_reference():
    # Appendix A2: the kernel result is narrowed int32 -> int16, values outside +-32767 wrap
    L = 6000
    q = np.full((1, L + 1), ord("A"), np.uint8); q[:, L] = 10
    s = np.full((2, L + 1), ord("C"), np.uint8); s[:, L] = 10
    s[1, :L] = ord("A")
    dp = R.dp_scores("nw", q, s, 4, -6, -10)
    assert dp.tolist() == [[-36000, 24000]]
    got = R.oracle_batch(R.ALGO_BITPAL_PACKED, q, s, M=4, I=-6, G=-10)
    assert (got == dp.astype(np.int16)).all() and got[0, 0] == 29536


def _banded_w(lib, qc, L, sp, e, wb):
    return lib.oracle_banded_myers_w(qc.ctypes.data, L, sp, L, e, wb)


def test_banded_word_width_and_padding_independence():
    """Appendix A8: as long as band+1 bits fit, the word width does not change the result; and the
    bytes the reference reads past the subject end never reach the cells that are read out."""
    lib = R.oracle()
    lib.oracle_banded_myers_w.restype = C.c_int8
    lib.oracle_banded_myers_w.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    rng = np.random.default_rng(5)
    for _ in range(40):
        L = int(rng.integers(8, 700)); e = int(rng.integers(1, 16))
        if L < 2 * e + 2:
            continue
        q = R.random_rows(rng, 1, L)
        s = np.concatenate([R.mutate_rows(rng, q[0, :L], 6, 2 * e + 2), R.indel_rows(rng, q[0, :L], 6, e + 3), R.random_rows(rng, 3, L)])
        qc = R.to_codes(q)
        for pad in (10, ord("A"), ord("T")):
            buf = np.concatenate([s.reshape(-1), np.full(64, pad, np.uint8)])
            for i in range(s.shape[0]):
                sp = buf.ctypes.data + i * (L + 1)
                r64 = _banded_w(lib, qc, L, sp, e, 64)
                assert _banded_w(lib, qc, L, sp, e, 2 * e + 2) == r64
                assert _banded_w(lib, qc, L, sp, e, 32) == r64
        # last row with different padding bytes gives the same value
        i = s.shape[0] - 1
        vals = set()
        for pad in (10, ord("A"), ord("G")):
            buf = np.concatenate([s.reshape(-1), np.full(64, pad, np.uint8)])
            vals.add(_banded_w(lib, qc, L, buf.ctypes.data + i * (L + 1), e, 64))
        assert len(vals) == 1


# ---- live against the unmodified reference (only where oracle/_ref travelled / was built) -------
LIVE = [("myers_cpu", R.ALGO_MYERS_GLOBAL), ("myers_sse", R.ALGO_MYERS_GLOBAL), ("bitpal_avx2", R.ALGO_BITPAL_PACKED),
        ("bitpal_avx512", R.ALGO_BITPAL_PACKED), ("semiglobal_cpu", R.ALGO_MYERS_SEMIGLOBAL)]


@pytest.mark.parametrize("variant,algo", LIVE)
def test_oracle_equals_reference_live(variant, algo):
    if not R.ref_available(variant):
        pytest.skip(f"oracle/_ref/libref_{variant}.so not available on this host")
    rng = np.random.default_rng(21)
    ref = R.reflib(variant)
    for _ in range(6):
        ql, sl = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        q = R.random_rows(rng, 2, ql, with_n=0.03)
        s = R.random_rows(rng, 41, sl, with_n=0.03)
        s[:10, : min(ql, sl)] = q[0, : min(ql, sl)]
        assert (ref.run(q, s) == R.oracle_batch(algo, q, s)).all(), (variant, ql, sl)


def test_banded_oracle_equals_reference_live():
    if not R.ref_available("banded_cpu"):
        pytest.skip("oracle/_ref/libref_banded_cpu.so not available on this host")
    rng = np.random.default_rng(22)
    ref = R.reflib("banded_cpu")
    checked = 0
    for _ in range(40):
        L = int(rng.integers(8, 400)); e = int(rng.integers(1, min(31, L // 2) + 1))
        # the reference's own Peq builder overflows its buffer outside this predicate (banded_host.h)
        if not ((L - 1) // 64 + 1 < (L - e + 63) // 64 + 1):
            continue
        q = R.random_rows(rng, 1, L)
        s = np.concatenate([R.mutate_rows(rng, q[0, :L], 30, 2 * e + 2), R.indel_rows(rng, q[0, :L], 30, e + 3), R.random_rows(rng, 10, L)])
        assert (ref.run(q, s, e=e) == R.oracle_batch(R.ALGO_BANDED, q, s, e=e)).all(), (L, e)
        checked += 1
    assert checked >= 10
