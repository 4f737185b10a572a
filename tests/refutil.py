"""Test-side helpers: synthetic data, the CPU oracle (oracle/liboracle.so) and the unmodified
reference compiled into oracle/_ref/ (called in-process through ctypes).

TEST INFRASTRUCTURE ONLY -- nothing in bgsa_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
REF_DIR = ORACLE_DIR / "_ref"

ALGO_MYERS_GLOBAL, ALGO_MYERS_SEMIGLOBAL, ALGO_BANDED, ALGO_BITPAL_PACKED, ALGO_BITPAL_NONPACKED, ALGO_BITPAL_SEMI = range(6)

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_MAP = np.zeros(256, dtype=np.uint8)
for _i, _c in enumerate(b"ACGTN"):
    _MAP[_c] = _i


# ------------------------------------------------------------------------------------------
# data
# ------------------------------------------------------------------------------------------
def random_rows(rng: np.random.Generator, count: int, length: int, with_n: float = 0.0) -> np.ndarray:
    """[count, length+1] uint8 ASCII rows terminated by '\\n' (seq_t.content layout, global.h:9-16)."""
    rows = np.empty((count, length + 1), dtype=np.uint8)
    rows[:, :length] = _ACGT[rng.integers(0, 4, size=(count, length))]
    if with_n > 0:
        mask = rng.random((count, length)) < with_n
        rows[:, :length][mask] = ord("N")
    rows[:, length] = ord("\n")
    return rows


def mutate_rows(rng: np.random.Generator, base: np.ndarray, count: int, max_subs: int) -> np.ndarray:
    """`count` copies of base (1-D ASCII, no newline) with k in U{0..max_subs} random substitutions."""
    length = base.shape[0]
    rows = np.empty((count, length + 1), dtype=np.uint8)
    rows[:, :length] = base
    rows[:, length] = ord("\n")
    k = rng.integers(0, max_subs + 1, size=count)
    for j in range(max_subs):
        sel = np.nonzero(k > j)[0]
        pos = rng.integers(0, length, size=sel.shape[0])
        rows[sel, pos] = _ACGT[rng.integers(0, 4, size=sel.shape[0])]
    return rows


def indel_rows(rng: np.random.Generator, base: np.ndarray, count: int, max_edits: int) -> np.ndarray:
    """copies of base with up to max_edits random sub/ins/del, re-trimmed/padded to len(base)."""
    length = base.shape[0]
    rows = np.empty((count, length + 1), dtype=np.uint8)
    rows[:, length] = ord("\n")
    for i in range(count):
        s = list(base)
        for _ in range(int(rng.integers(0, max_edits + 1))):
            op = int(rng.integers(0, 3)); p = int(rng.integers(0, len(s)))
            if op == 0:
                s[p] = int(_ACGT[rng.integers(0, 4)])
            elif op == 1:
                s.insert(p, int(_ACGT[rng.integers(0, 4)]))
            elif len(s) > 1:
                del s[p]
        while len(s) < length:
            s.append(int(_ACGT[rng.integers(0, 4)]))
        rows[i, :length] = s[:length]
    return rows


def to_codes(rows: np.ndarray) -> np.ndarray:
    """file.c:135-139: map every byte except '\\n' through mapping_table."""
    out = _MAP[rows]
    out[rows == ord("\n")] = ord("\n")
    return out


def write_rows(path, rows: np.ndarray, final_newline: bool = True) -> None:
    data = rows.tobytes()
    if not final_newline:
        data = data[:-1]
    Path(path).write_bytes(data)


def read_rows(path) -> np.ndarray:
    data = Path(path).read_bytes()
    if not data.endswith(b"\n"):
        data += b"\n"
    length = data.index(b"\n")
    arr = np.frombuffer(data, dtype=np.uint8)
    return arr.reshape(-1, length + 1).copy()


def numpy_pack(rows, layout):
    """rows [n, slen+1] ASCII -> (codes uint32 [ntiles, ku, 32, 4], nmask uint32 [ntiles, kn, 32], has_n [ntiles])."""
    n, slen = rows.shape[0], rows.shape[1] - 1
    ntiles, ku, kn = (n + 31) // 32, (slen + 63) // 64, (slen + 31) // 32
    code = np.zeros(256, dtype=np.uint64)
    code[ord("C")], code[ord("G")], code[ord("T")] = 1, 2, 3
    c = np.zeros((ntiles * 32, ku * 64), dtype=np.uint64)
    c[:n, :slen] = code[rows[:, :slen]]
    isn = np.zeros((ntiles * 32, kn * 32), dtype=np.uint64)
    isn[:n, :slen] = rows[:, :slen] == ord("N")
    if layout == 0:      # base i of a unit at bits 2*(i%16) of word i/16
        w = (c.reshape(ntiles, 32, ku, 4, 16) << (2 * np.arange(16, dtype=np.uint64))).sum(-1)
    else:                # x,y = low/high planes of bases 0..31, z,w = of bases 32..63
        b = c.reshape(ntiles, 32, ku, 2, 32)
        lo = ((b & 1) << np.arange(32, dtype=np.uint64)).sum(-1)
        hi = ((b >> 1) << np.arange(32, dtype=np.uint64)).sum(-1)
        w = np.stack([lo[..., 0], hi[..., 0], lo[..., 1], hi[..., 1]], axis=-1)
    codes = w.transpose(0, 2, 1, 3).astype(np.uint32)
    nm = (isn.reshape(ntiles, 32, kn, 32) << np.arange(32, dtype=np.uint64)).sum(-1).transpose(0, 2, 1).astype(np.uint32)
    return codes, nm, isn.reshape(ntiles, -1).any(axis=1)


def split_packed(raw: np.ndarray, slen: int, n: int):
    """The three regions of a packed buffer (csrc/bgsa_common.cuh make_packed_view): codes, N plane, per-tile flags."""
    ntiles, ku, kn = (n + 31) // 32, (slen + 63) // 64, (slen + 31) // 32
    up = lambda x: (x + 255) // 256 * 256      # noqa: E731
    codes = raw[: ntiles * ku * 512].view(np.uint32).reshape(ntiles, ku, 32, 4)
    o1 = up(ntiles * ku * 512)
    nm = raw[o1: o1 + ntiles * kn * 128].view(np.uint32).reshape(ntiles, kn, 32)
    flags = raw[o1 + up(ntiles * kn * 128): o1 + up(ntiles * kn * 128) + ntiles]
    return codes, nm, flags


# ------------------------------------------------------------------------------------------
# oracle (our restatement)
# ------------------------------------------------------------------------------------------
_oracle = None


def build_oracle() -> Path:
    so = ORACLE_DIR / "liboracle.so"
    src = ORACLE_DIR / "bgsa_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(ORACLE_DIR), "oracle", "GCC=gcc"])
    return so


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(str(build_oracle()))
        lib.oracle_align_batch.restype = C.c_int
        lib.oracle_align_batch.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                                           C.c_int, C.c_void_p, C.c_int]
        for name in ("oracle_dp_edit", "oracle_dp_semiglobal"):
            f = getattr(lib, name); f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        for name in ("oracle_dp_nw", "oracle_dp_nw_semiglobal"):
            f = getattr(lib, name); f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        _oracle = lib
    return _oracle


def oracle_batch(algo: int, queries: np.ndarray, subjects: np.ndarray, M=2, I=-3, G=-5, e=5, threads=0) -> np.ndarray:
    """queries/subjects: [n, len+1] uint8 ASCII rows.  Returns [nq, ns] int16 (int8 for banded)."""
    lib = oracle()
    q = np.ascontiguousarray(to_codes(queries))
    pad = np.zeros(64, dtype=np.uint8) + ord("\n")       # banded Peq builder reads e bytes past the end
    s = np.concatenate([np.ascontiguousarray(subjects).reshape(-1), pad])
    nq, qlen = q.shape[0], q.shape[1] - 1
    ns, slen = subjects.shape[0], subjects.shape[1] - 1
    out = np.zeros((nq, ns), dtype=np.int8 if algo == ALGO_BANDED else np.int16)
    rc = lib.oracle_align_batch(algo, M, I, G, e, q.ctypes.data, nq, qlen, s.ctypes.data, ns, slen,
                                out.ctypes.data, threads)
    if rc != 0:
        raise ValueError("oracle_align_batch rejected its arguments")
    return out


def dp_scores(kind: str, queries: np.ndarray, subjects: np.ndarray, M=2, I=-3, G=-5) -> np.ndarray:
    lib = oracle()
    q = np.ascontiguousarray(to_codes(queries)); s = np.ascontiguousarray(subjects)
    nq, qlen = q.shape[0], q.shape[1] - 1
    ns, slen = s.shape[0], s.shape[1] - 1
    out = np.zeros((nq, ns), dtype=np.int64)
    for i in range(nq):
        for j in range(ns):
            qa, sa = q[i].ctypes.data, s[j].ctypes.data
            if kind == "edit":
                out[i, j] = lib.oracle_dp_edit(qa, qlen, sa, slen)
            elif kind == "semi":
                out[i, j] = lib.oracle_dp_semiglobal(qa, qlen, sa, slen)
            elif kind == "nw_semi":
                out[i, j] = lib.oracle_dp_nw_semiglobal(qa, qlen, sa, slen, M, I, G)
            else:
                out[i, j] = lib.oracle_dp_nw(qa, qlen, sa, slen, M, I, G)
    return out


# ------------------------------------------------------------------------------------------
# the unmodified reference, in-process (oracle/_ref/libref_*.so)
# ------------------------------------------------------------------------------------------
class SeqT(C.Structure):  # original/BGSA_CPU/global.h:9-16
    _fields_ = [("len", C.c_int), ("size", C.c_int64), ("count", C.c_int64), ("extra_size", C.c_int),
                ("extra_count", C.c_int), ("content", C.c_void_p)]


_VARIANTS = {
    #  name            prefix  V   wordbits vec_bytes  result   needs
    "myers_cpu":      ("cpu", 1, 64, 8, np.int16, None),
    "semiglobal_cpu": ("cpu", 1, 64, 8, np.int16, None),
    "myers_sse":      ("sse", 4, 32, 16, np.int16, "sse4_1"),
    "bitpal_avx2":    ("avx", 8, 32, 32, np.int16, "avx2"),
    "bitpal_avx512":  ("mic", 16, 32, 64, np.int16, "avx512f"),
    "banded_cpu":     ("cpu", 1, 64, 8, np.int8, None),
}


def cpu_flags() -> set:
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def ref_available(variant: str) -> bool:
    need = _VARIANTS[variant][5]
    return (REF_DIR / f"libref_{variant}.so").exists() and (need is None or need in cpu_flags())


def _aligned(nbytes: int, align: int = 64) -> np.ndarray:
    raw = np.zeros(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


class RefLib:
    """Calls <arch>_handle_reads + <arch>_cal_align_score of one reference variant
    (original/BGSA_CPU/global.c:25-70, cal_cpu.c:43-85 and their SIMD twins)."""

    def __init__(self, variant: str, threads: int = 0):
        prefix, self.V, self.wordbits, self.vec_bytes, self.rtype, _ = _VARIANTS[variant]
        self.variant = variant
        self.lib = C.CDLL(str(REF_DIR / f"libref_{variant}.so"), mode=os.RTLD_LOCAL)
        self.handle = getattr(self.lib, f"{prefix}_handle_reads")
        self.handle.restype = None
        self.handle.argtypes = [C.POINTER(SeqT), C.c_void_p, C.c_int, C.c_int64, C.c_int64]
        self.cal = getattr(self.lib, f"{prefix}_cal_align_score")
        self.cal.restype = None
        self.cal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]
        self.lib.init_mapping_table()
        self.threads = threads or os.cpu_count() or 1
        C.c_int.in_dll(self.lib, "cpu_threads").value = self.threads
        self.dvdh_len = C.c_int.in_dll(self.lib, "dvdh_len").value
        self.full_bits = C.c_int.in_dll(self.lib, "full_bits").value
        self.banded = variant.startswith("banded")

    def geometry(self, qlen: int, slen: int, e: int):
        if self.banded:
            h = e + slen - qlen
            word_num = (slen - h + 63) // 64 + 1            # banded/BGSA_CPU/cal_cpu.c:253-254
        elif self.full_bits:
            word_num = (slen + self.wordbits - 1) // self.wordbits
        else:
            word_num = (slen + self.wordbits - 2) // (self.wordbits - 1)
        chunk = (4000 + slen - 1) // slen                   # main.c:89, cal_cpu.c:257
        return word_num, chunk

    def prepare(self, queries: np.ndarray, subjects: np.ndarray, e: int = 5):
        """Pads to V_NUM (file.c:96-111), allocates Peq/result/scratch like cal_on_*()."""
        if self.banded:
            C.c_int.in_dll(self.lib, "threshold").value = e
        nq, qlen = queries.shape[0], queries.shape[1] - 1
        ns, slen = subjects.shape[0], subjects.shape[1] - 1
        pad = (-ns) % self.V
        content = _aligned((ns + pad) * (slen + 1) + 128)
        content[: ns * (slen + 1)] = subjects.reshape(-1)
        if pad:
            tail = np.full((pad, slen + 1), ord("N"), dtype=np.uint8); tail[:, slen] = ord("\n")
            content[ns * (slen + 1):(ns + pad) * (slen + 1)] = tail.reshape(-1)
        content[(ns + pad) * (slen + 1):] = ord("\n")
        seq = SeqT(slen, (ns + pad) * (slen + 1), ns + pad, 0, pad, content.ctypes.data)
        word_num, chunk = self.geometry(qlen, slen, e)
        peq = _aligned((self.wordbits // 8) * word_num * 5 * (ns + pad) + 64)
        per_thread = self.dvdh_len if self.banded else word_num * self.dvdh_len
        scratch = _aligned(self.vec_bytes * per_thread * self.threads + 64)
        results = _aligned(np.dtype(self.rtype).itemsize * nq * (ns + pad) + 64)
        qcodes = np.ascontiguousarray(to_codes(queries))
        return dict(nq=nq, qlen=qlen, ns=ns, slen=slen, pad=pad, content=content, seq=seq, word_num=word_num,
                    chunk=chunk, peq=peq, scratch=scratch, results=results, qcodes=qcodes)

    def handle_reads(self, st) -> None:
        st["peq"][:] = 0                                    # cal_cpu.c:273 memset
        self.handle(C.byref(st["seq"]), st["peq"].ctypes.data, st["word_num"], 0, st["ns"] + st["pad"])

    def cal_align_score(self, st) -> np.ndarray:
        self.cal(st["qcodes"].ctypes.data, st["peq"].ctypes.data, st["results"].ctypes.data, st["qlen"], st["nq"],
                 st["slen"], st["ns"] + st["pad"], 0, st["nq"], st["word_num"], st["chunk"], st["scratch"].ctypes.data)
        n = st["nq"] * (st["ns"] + st["pad"])
        res = st["results"][: n * np.dtype(self.rtype).itemsize].view(self.rtype).reshape(st["nq"], st["ns"] + st["pad"])
        return res[:, : st["ns"]].copy()

    def run(self, queries: np.ndarray, subjects: np.ndarray, e: int = 5) -> np.ndarray:
        st = self.prepare(queries, subjects, e)
        self.handle_reads(st)
        if self.banded:
            # banded early exit writes nothing but 127 for tripped subjects and leaves
            # non-tripped untouched before the final store; the result buffer needs no preset.
            pass
        return self.cal_align_score(st)


_reflibs: dict = {}


def reflib(variant: str) -> RefLib:
    if variant not in _reflibs:
        _reflibs[variant] = RefLib(variant)
    return _reflibs[variant]


def sample_data():
    """The reference's only fixture (3 x 500 bp queries, 128 x 500 bp subjects), committed as
    tests/golden/sample_{query,subject}.txt by tests/golden/make_golden.py."""
    g = ROOT / "tests" / "golden"
    return read_rows(g / "sample_query.txt"), read_rows(g / "sample_subject.txt")
