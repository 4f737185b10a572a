"""e2e tuning: time bgsa_align_batch (pinned host buffers) for a workload; chunk count comes from BGSA_CHUNKS."""
import sys, time, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import numpy as np, torch
import bgsa_b200 as B, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
algo = {"C2": 3, "C3": 2, "C4": 1, "myers150": 0}[name]
cfgname = "C2" if name == "myers150" else name
count = int(sys.argv[2]) if len(sys.argv) > 2 else synth.CONFIGS[cfgname]["count"]
q, s = synth.make(cfgname, count)
kw = {"threshold": 5} if algo == 2 else {}
p = B.Params.default(algo, **kw)
h = torch.from_numpy(s.reshape(-1)).pin_memory()
sp = h.numpy().reshape(s.shape)
out = torch.empty(s.shape[0] * 2, dtype=torch.uint8).pin_memory().numpy().view(np.int8 if algo == 2 else np.int16)[: s.shape[0]].reshape(1, -1)
for _ in range(3): B.align_batch(p, q, sp, out=out)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 10
for _ in range(n): B.align_batch(p, q, sp, out=out)
t = (time.perf_counter() - t0) / n
cells = (q.shape[1]-1) * (s.shape[1]-1) * s.shape[0]
# raw H2D for comparison
d = torch.empty_like(h, device="cuda")
torch.cuda.synchronize(); t1 = time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); th = (time.perf_counter() - t1) / 5
print(f"{name} chunks={os.environ.get("BGSA_CHUNKS","default")} e2e {t*1e3:.3f} ms -> {cells/t/1e9:.0f} GCUPS ; raw H2D {th*1e3:.3f} ms ({h.numel()/th/1e9:.1f} GB/s)")
