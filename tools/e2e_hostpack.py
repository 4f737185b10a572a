"""End-to-end A/B of the batch entry's two front ends (device pack of ASCII rows shipped over PCIe vs host-thread pack,
csrc/host_pack.cpp): bgsa_align_batch on pinned and on pageable subject memory, BGSA_HOST_PACK = 0 / 1 / auto.
    python tools/e2e_hostpack.py [workload,...]"""
import os, sys, time
from pathlib import Path
R = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(R), str(R / "tools"), str(R / "tests")]
import numpy as np, torch
import bgsa_b200 as B, synth

W = {"C2": ("C2", 3, 1_000_000, {}), "C3": ("C3", 2, 10_000_000, {"threshold": 5}), "C3s": ("C3s", 2, 10_000_000, {"threshold": 5}),
     "myers150": ("C2", 0, 1_000_000, {}), "C4": ("C4", 1, 300_000, {}), "C2np": ("C2", 4, 1_000_000, {})}
only = sys.argv[1].split(",") if len(sys.argv) > 1 else ["C2", "C3", "myers150", "C4"]
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "2", "auto"]
mems = sys.argv[3].split(",") if len(sys.argv) > 3 else ["pinned", "pageable"]
print("host pack pool:", B.host_pack_info(), "cores", os.cpu_count())
for name in only:
    cfg, algo, n, kw = W[name]
    q, s = synth.make(cfg, n)
    p = B.Params.default(algo, **kw)
    h = torch.from_numpy(s.reshape(-1)).pin_memory()
    sp = h.numpy().reshape(s.shape)
    es = 1 if algo == B.BANDED_MYERS else 2
    hr = torch.empty(n * es, dtype=torch.uint8).pin_memory()
    out = hr.numpy().view(np.int8 if es == 1 else np.int16).reshape(1, n)
    ref = None
    for mem, subj in (("pinned", sp), ("pageable", s)):
        if mem not in mems:
            continue
        for mode in modes:
            if mode == "auto":
                os.environ.pop("BGSA_HOST_PACK", None)
            else:
                os.environ["BGSA_HOST_PACK"] = mode
            for _ in range(3):
                B.align_batch(p, q, subj, out=out)
            t = time.perf_counter()
            reps = 10
            for _ in range(reps):
                B.align_batch(p, q, subj, out=out)
            dt = (time.perf_counter() - t) / reps
            if ref is None:
                ref = out.copy()
            ok = (out == ref).all()
            cells = (q.shape[1] - 1) * (s.shape[1] - 1) * n
            print(f"{name:9s} {mem:8s} host_pack={mode:4s} {dt * 1e3:8.3f} ms  {cells / dt / 1e9:9.0f} GCUPS  rows {s.nbytes / dt / 1e9:6.1f} GB/s  same={ok}", flush=True)
