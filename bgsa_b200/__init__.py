"""bgsa_b200 -- Python host-side mirror of the C ABI in include/bgsa_b200.h.

The product is libbgsa_b200.so (hand-written sm_100a CUDA kernels + a C shim) and the C host
tools next to it; this module only binds the C ABI with ctypes so that the tests and bench.py can
drive it.  There is NO CPU fallback: importing works without a GPU (symbols can be inspected),
but every compute call raises BgsaError when CUDA is unavailable, and load() raises if the shared
library has not been built (``make lib`` / ``__graft_entry__.build()``).

Interface mirrored (reference file:line each call replaces is in include/bgsa_b200.h):
    align_batch(...)          <-> cpu_cal_align_score + cpu_handle_reads  (cal.h:48, global.h:24)
    pack_subjects_device(...) <-> cpu_handle_reads                         (global.c:25-70)
    align_device(...)         <-> align_cpu / align_sse / align_avx / align_mic, batched
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

__all__ = [
    "MYERS_GLOBAL", "MYERS_SEMIGLOBAL", "BANDED_MYERS", "BITPAL_PACKED", "BITPAL_NONPACKED", "BITPAL_PACKED_SEMIGLOBAL",
    "BgsaError", "Params", "SeqT", "load", "lib_path", "align_batch", "result_dtype", "to_codes",
    "align_batch_submit", "align_batch_wait", "init_devices",
    "packed_bytes", "pack_subjects_device", "pack_subjects_host", "host_pack_info", "align_device", "align_rows_device", "int_peak", "bind_thread_to_device", "launch_count", "kernel_name", "rows_kernel_name", "supported", "jit_precompile", "batch_front_end",
]

MYERS_GLOBAL, MYERS_SEMIGLOBAL, BANDED_MYERS, BITPAL_PACKED, BITPAL_NONPACKED, BITPAL_PACKED_SEMIGLOBAL = range(6)
_STATUS = {1: "BGSA_ERR_ARG", 2: "BGSA_ERR_UNSUPPORTED", 3: "BGSA_ERR_CUDA", 4: "BGSA_ERR_NOMEM"}


class BgsaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{_STATUS.get(code, code)}: {message}")
        self.code = code


class Params(C.Structure):  # bgsa_params_t
    _fields_ = [("algo", C.c_int32), ("match", C.c_int32), ("mismatch", C.c_int32), ("gap", C.c_int32),
                ("threshold", C.c_int32), ("myers_sign", C.c_int32)]

    @classmethod
    def default(cls, algo: int, **kw) -> "Params":
        p = cls()
        load().bgsa_params_default(C.byref(p), algo)
        for k, v in kw.items():
            setattr(p, k, v)
        return p


class SeqT(C.Structure):  # bgsa_seq_t == seq_t (original/BGSA_CPU/global.h:9-16)
    _fields_ = [("len", C.c_int32), ("size", C.c_int64), ("count", C.c_int64), ("extra_size", C.c_int32),
                ("extra_count", C.c_int32), ("content", C.c_void_p)]


_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("BGSA_B200_LIB", Path(__file__).resolve().parent / "libbgsa_b200.so"))


def load():
    """Loads libbgsa_b200.so (in-tree).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise ImportError(f"{path} not built: run `make lib` (or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(str(path))
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    PP = C.POINTER(Params)
    sig = {
        "bgsa_version": (C.c_char_p, []),
        "bgsa_last_error": (C.c_char_p, []),
        "bgsa_device_count": (i32, [C.POINTER(i32)]),
        "bgsa_init_devices": (i32, [i32]),
        "bgsa_params_default": (None, [PP, i32]),
        "bgsa_result_size": (i32, [i32]),
        "bgsa_supported": (i32, [PP, i32, i32]),
        "bgsa_align_batch": (i32, [PP, vp, i32, i32, C.POINTER(SeqT), i64, i64, vp, i64, i32]),
        "bgsa_align_batch_submit": (i32, [PP, vp, i32, i32, C.POINTER(SeqT), i64, i64, vp, i64, i32, i32]),
        "bgsa_align_batch_wait": (i32, [i32, i32]),
        "bgsa_malloc_host": (vp, [C.c_size_t]),
        "bgsa_free_host": (None, [vp]),
        "bgsa_host_register": (i32, [vp, C.c_size_t]),
        "bgsa_host_unregister": (i32, [vp]),
        "bgsa_bind_thread_to_device": (i32, [i32, C.POINTER(i32)]),
        "bgsa_packed_bytes": (i64, [i32, i64]),
        "bgsa_pack_subjects_device": (i32, [PP, vp, i32, i64, vp, i32, vp]),
        "bgsa_pack_subjects_host": (i32, [PP, vp, i32, i64, vp]),
        "bgsa_host_pack_info": (i32, [C.POINTER(i32), C.c_char_p, i32]),
        "bgsa_align_device": (i32, [PP, vp, i32, i32, vp, i32, i64, vp, i64, i32, vp]),
        "bgsa_align_rows_device": (i32, [PP, vp, i32, i32, vp, i32, i64, vp, i64, i32, vp]),
        "bgsa_launch_count": (i64, []),
        "bgsa_kernel_name": (i32, [PP, i32, i32, C.c_char_p, i32]),
        "bgsa_rows_kernel_name": (i32, [PP, i32, i32, C.c_char_p, i32, C.POINTER(i32)]),
        "bgsa_jit_precompile": (i32, [PP, i32, i32]),
        "bgsa_batch_front_end": (i32, [i32, i32, C.POINTER(C.c_double)]),
        "bgsa_int_peak": (i32, [i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "bgsa_align_peq_chunk": (i32, [PP, vp, i32, vp, i32, i32, i32, i32, i32, i64, vp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)       # AttributeError here = the library does not export the ABI
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "bgsa_version", "bgsa_last_error", "bgsa_device_count", "bgsa_init_devices", "bgsa_params_default", "bgsa_result_size", "bgsa_supported",
    "bgsa_align_batch", "bgsa_align_batch_submit", "bgsa_align_batch_wait", "bgsa_malloc_host", "bgsa_free_host",
    "bgsa_host_register", "bgsa_host_unregister", "bgsa_bind_thread_to_device",
    "bgsa_packed_bytes", "bgsa_pack_subjects_device", "bgsa_pack_subjects_host", "bgsa_host_pack_info", "bgsa_align_device", "bgsa_align_rows_device", "bgsa_launch_count", "bgsa_kernel_name", "bgsa_rows_kernel_name", "bgsa_jit_precompile", "bgsa_batch_front_end",
    "bgsa_int_peak", "bgsa_align_peq_chunk",
]


def _check(rc: int) -> None:
    if rc != 0:
        raise BgsaError(rc, load().bgsa_last_error().decode())


_MAP = np.zeros(256, dtype=np.uint8)
for _i, _c in enumerate(b"ACGTN"):
    _MAP[_c] = _i


def to_codes(rows: np.ndarray) -> np.ndarray:
    """ASCII query rows -> the 0..4 codes get_ref_from_file() produces (file.c:135-139)."""
    out = _MAP[rows]
    out[rows == 10] = 10
    return out


def result_dtype(algo: int):
    return np.int8 if algo == BANDED_MYERS else np.int16


def supported(params: Params, query_len: int, subject_len: int) -> bool:
    return load().bgsa_supported(C.byref(params), query_len, subject_len) == 0


def kernel_name(params: Params, query_len: int, subject_len: int) -> str:
    buf = C.create_string_buffer(160)
    _check(load().bgsa_kernel_name(C.byref(params), query_len, subject_len, buf, 160))
    return buf.value.decode()


def rows_kernel_name(params: Params, query_len: int, subject_len: int):
    """(name, fused): what align_rows_device runs -- fused = ONE kernel fed with the ASCII rows (no pack launch)."""
    buf = C.create_string_buffer(256)
    fused = C.c_int(0)
    _check(load().bgsa_rows_kernel_name(C.byref(params), query_len, subject_len, buf, 256, C.byref(fused)))
    return buf.value.decode(), fused.value == 1


def jit_precompile(params: Params, query_len: int, subject_len: int) -> None:
    """Compiles (NVRTC) and caches the kernels of a scoring scheme that is not built into the library; needs no GPU."""
    _check(load().bgsa_jit_precompile(C.byref(params), query_len, subject_len))


def batch_front_end(device: int = 0, slot: int = 0) -> float:
    """Share of host-packed chunks of the last batch job on (device, slot): 0 = all ASCII over the link, 1 = all packed on the host."""
    v = C.c_double(0.0)
    _check(load().bgsa_batch_front_end(device, slot, C.byref(v)))
    return v.value


def launch_count() -> int:
    return int(load().bgsa_launch_count())


def make_seq(subjects: np.ndarray) -> SeqT:
    """subjects: C-contiguous [count, len+1] uint8 ASCII rows (each ending in '\\n')."""
    assert subjects.dtype == np.uint8 and subjects.ndim == 2 and subjects.flags.c_contiguous
    count, stride = subjects.shape
    return SeqT(stride - 1, count * stride, count, 0, 0, subjects.ctypes.data)


def align_batch(params: Params, queries: np.ndarray, subjects: np.ndarray, first: int = 0, count: int | None = None,
                device: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    """Host-buffer entry.  queries: [nq, qlen+1] ASCII rows; subjects: [ns, slen+1] ASCII rows.
    Returns [nq, count] scores (int16, or int8 for BANDED_MYERS)."""
    lib = load()
    qc = np.ascontiguousarray(to_codes(np.asarray(queries, dtype=np.uint8)))
    seq = make_seq(subjects)
    if count is None:
        count = subjects.shape[0] - first
    nq, qlen = qc.shape[0], qc.shape[1] - 1
    if out is None:
        out = np.zeros((nq, count), dtype=result_dtype(params.algo))
    _check(lib.bgsa_align_batch(C.byref(params), qc.ctypes.data, nq, qlen, C.byref(seq), first, count,
                                out.ctypes.data, out.strides[0] // out.itemsize, device))
    return out


def init_devices(n_devices: int) -> None:
    """Creates the contexts (streams, events) of devices 0..n-1 in parallel (bgsa_init_devices)."""
    _check(load().bgsa_init_devices(n_devices))


def align_batch_submit(params: Params, queries: np.ndarray, subjects: np.ndarray, first: int, count: int, out: np.ndarray,
                       device: int = 0, slot: int = 0) -> None:
    """Asynchronous half of align_batch (bgsa_align_batch_submit): enqueues H2D + kernels + D2H of subjects
    [first, first+count) on `device` and returns.  `subjects` and `out` must stay alive (and should be pinned) until
    align_batch_wait(device, slot).  out: [nq, >= count] scores, written at column 0.. -- the host-side mirror of one
    device's share in cal_mic.c:459-481."""
    qc = np.ascontiguousarray(to_codes(np.asarray(queries, dtype=np.uint8)))
    seq = make_seq(subjects)
    nq, qlen = qc.shape[0], qc.shape[1] - 1
    assert out.dtype == result_dtype(params.algo) and out.shape[0] == nq and out.shape[1] >= count
    _check(load().bgsa_align_batch_submit(C.byref(params), qc.ctypes.data, nq, qlen, C.byref(seq), first, count,
                                          out.ctypes.data, out.strides[0] // out.itemsize, device, slot))


def align_batch_wait(device: int = 0, slot: int = 0) -> None:
    _check(load().bgsa_align_batch_wait(device, slot))


def packed_bytes(subject_len: int, count: int) -> int:
    return int(load().bgsa_packed_bytes(subject_len, count))


def pack_subjects_device(params: Params, d_rows_ptr: int, subject_len: int, count: int, d_packed_ptr: int,
                         device: int = 0, stream: int = 0) -> None:
    _check(load().bgsa_pack_subjects_device(C.byref(params), d_rows_ptr, subject_len, count, d_packed_ptr, device, stream))


def pack_subjects_host(params: Params, subjects: np.ndarray) -> np.ndarray:
    """ASCII rows -> packed tiles on the host cores (bgsa_pack_subjects_host); returns the packed bytes."""
    assert subjects.dtype == np.uint8 and subjects.ndim == 2 and subjects.flags.c_contiguous
    count, slen = subjects.shape[0], subjects.shape[1] - 1
    out = np.zeros(packed_bytes(slen, count), dtype=np.uint8)
    _check(load().bgsa_pack_subjects_host(C.byref(params), subjects.ctypes.data, slen, count, out.ctypes.data))
    return out


def host_pack_info():
    t = C.c_int(0)
    buf = C.create_string_buffer(32)
    _check(load().bgsa_host_pack_info(C.byref(t), buf, 32))
    return t.value, buf.value.decode()


def align_device(params: Params, queries: np.ndarray, d_packed_ptr: int, subject_len: int, count: int,
                 d_results_ptr: int, result_stride: int, device: int = 0, stream: int = 0) -> None:
    qc = np.ascontiguousarray(to_codes(np.asarray(queries, dtype=np.uint8)))
    _check(load().bgsa_align_device(C.byref(params), qc.ctypes.data, qc.shape[0], qc.shape[1] - 1, d_packed_ptr,
                                    subject_len, count, d_results_ptr, result_stride, device, stream))


def bind_thread_to_device(device: int = 0) -> int:
    """Pins the calling thread to the CPUs of the GPU's NUMA node; returns the node (-1: none exposed)."""
    node = C.c_int(-1)
    _check(load().bgsa_bind_thread_to_device(device, C.byref(node)))
    return node.value


def align_rows_device(params: Params, queries: np.ndarray, d_rows_ptr: int, subject_len: int, count: int,
                      d_results_ptr: int, result_stride: int, device: int = 0, stream: int = 0) -> None:
    """Device-resident ASCII rows -> scores (pack + align, or the fused banded kernel)."""
    qc = np.ascontiguousarray(to_codes(np.asarray(queries, dtype=np.uint8)))
    _check(load().bgsa_align_rows_device(C.byref(params), qc.ctypes.data, qc.shape[0], qc.shape[1] - 1, d_rows_ptr,
                                         subject_len, count, d_results_ptr, result_stride, device, stream))


def int_peak(device: int = 0):
    ops, mhz = C.c_double(), C.c_double()
    _check(load().bgsa_int_peak(device, C.byref(ops), C.byref(mhz)))
    return ops.value, mhz.value
