"""Runs pack + align of one workload a few times (device-resident) -- the target of ncu captures:
    ncu --set full --import-source on -k regex:<kernel> -c 1 python tools/prof_one.py C3 [count]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import numpy as np, torch
import bgsa_b200 as B, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
W = {"C2": ("C2", 3, {}), "C3": ("C3", 2, {"threshold": 5}), "C4": ("C4", 1, {}), "C5": ("C5", 3, {}), "myers150": ("C2", 0, {}), "C2np": ("C2", 4, {}), "myers5k": ("C5", 0, {}), "C3s": ("C3s", 2, {"threshold": 5})}
cfg, algo, kw = W[name]
count = int(sys.argv[2]) if len(sys.argv) > 2 else {"C2": 1_000_000, "C3": 10_000_000, "C4": 300_000, "C5": 8192, "myers150": 1_000_000, "C2np": 300_000, "myers5k": 16384, "C3s": 10_000_000}[name]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
q, s = synth.make(cfg, count)
p = B.Params.default(algo, **kw)
sl = s.shape[1] - 1
d_rows = torch.from_numpy(s.reshape(-1)).cuda()
d_packed = torch.empty(B.packed_bytes(sl, count), dtype=torch.uint8, device="cuda")
d_res = torch.zeros(count * 2, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(reps):      # the same resident step as bench.py times
    if B.rows_kernel_name(p, q.shape[1] - 1, sl)[1]:      # one kernel fed with the ASCII rows
        B.align_rows_device(p, q, d_rows.data_ptr(), sl, count, d_res.data_ptr(), count, 0, st)
    else:
        B.pack_subjects_device(p, d_rows.data_ptr(), sl, count, d_packed.data_ptr(), 0, st)
        B.align_device(p, q, d_packed.data_ptr(), sl, count, d_res.data_ptr(), count, 0, st)
torch.cuda.synchronize()
print("done", name, count)
