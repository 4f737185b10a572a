"""Small end-to-end run of every algorithm (host-buffer API + device-resident API) for compute-sanitizer:
    compute-sanitizer --tool memcheck python tests/scripts/sanitize_small.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgsa_b200 as B, refutil as R

rng = np.random.default_rng(4)
bad = 0
for ql, sl, ns in [(150, 150, 300), (100, 100, 1000), (37, 61, 45), (1000, 1000, 70), (2100, 333, 40), (5000, 5000, 36), (64, 64, 64)]:
    q = R.random_rows(rng, 2, ql, with_n=0.01); s = R.random_rows(rng, ns, sl, with_n=0.01)
    for algo, oalgo in ((B.MYERS_GLOBAL, 0), (B.MYERS_SEMIGLOBAL, 1), (B.BITPAL_PACKED, 3), (B.BITPAL_PACKED_SEMIGLOBAL, 5)):
        got = B.align_batch(B.Params.default(algo), q, s)
        bad += int((got != R.oracle_batch(oalgo, q, s)).sum())
    if ql <= 2100:
        bad += int((B.align_batch(B.Params.default(B.BITPAL_NONPACKED), q, s) != R.oracle_batch(3, q, s)).sum())
    if ql == sl:
        got = B.align_batch(B.Params.default(B.BANDED_MYERS, threshold=5), q, s)
        if ((ql - 1) // 64 + 1) < ((ql - 5 + 63) // 64 + 1):
            bad += int((got != R.oracle_batch(2, q, s, e=5)).sum())
# device-resident entries with an exactly-sized row buffer
q = R.random_rows(rng, 1, 150); s = R.random_rows(rng, 257, 150)
p = B.Params.default(B.BITPAL_PACKED)
d_rows = torch.from_numpy(s.reshape(-1)).cuda()
d_packed = torch.empty(B.packed_bytes(150, 257), dtype=torch.uint8, device="cuda")
d_res = torch.zeros(257 * 2, dtype=torch.uint8, device="cuda")
B.pack_subjects_device(p, d_rows.data_ptr(), 150, 257, d_packed.data_ptr())
B.align_device(p, q, d_packed.data_ptr(), 150, 257, d_res.data_ptr(), 257)
torch.cuda.synchronize()
bad += int((d_res.cpu().numpy().view(np.int16)[None, :] != R.oracle_batch(3, q, s)).sum())
print("mismatches", bad)
sys.exit(1 if bad else 0)
