"""First GPU contact: parity of every algorithm against the oracle on small inputs + rough timings."""
import sys, time, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import bgsa_b200 as B
import refutil as R

rng = np.random.default_rng(11)
print(B.load().bgsa_version(), torch.cuda.get_device_name(0))
bad = 0
def check(name, algo, q, s, **kw):
    global bad
    p = B.Params.default(algo, **{k: v for k, v in kw.items() if k in ("match", "mismatch", "gap", "threshold")})
    got = B.align_batch(p, q, s)
    oalgo = {B.MYERS_GLOBAL: 0, B.MYERS_SEMIGLOBAL: 1, B.BANDED_MYERS: 2, B.BITPAL_PACKED: 3, B.BITPAL_NONPACKED: 3}[algo]
    exp = R.oracle_batch(oalgo, q, s, M=kw.get("match", 2), I=kw.get("mismatch", -3), G=kw.get("gap", -5), e=kw.get("threshold", 5))
    ok = (got == exp).all()
    if not ok:
        bad += 1
        idx = np.argwhere(got != exp)
        print("MISMATCH", name, q.shape, s.shape, kw, "n=", len(idx), idx[:3].tolist(), got[got != exp][:5], exp[got != exp][:5])
    return ok

q, s = R.sample_data()
for algo in (B.MYERS_GLOBAL, B.MYERS_SEMIGLOBAL, B.BITPAL_PACKED, B.BITPAL_NONPACKED):
    print("sample", algo, check("sample", algo, q, s))
for ql, sl, ns in [(150, 150, 1000), (37, 150, 333), (150, 37, 65), (500, 480, 200), (1000, 1000, 130), (1, 1, 33), (33, 70, 31), (1024, 64, 40), (1500, 300, 70), (2100, 500, 40), (5000, 700, 33)]:
    qq = R.random_rows(rng, 2, ql, with_n=0.01); ss = R.random_rows(rng, ns, sl, with_n=0.01)
    ss[: ns // 3, : min(ql, sl)] = qq[0, : min(ql, sl)]
    for algo in (B.MYERS_GLOBAL, B.MYERS_SEMIGLOBAL, B.BITPAL_PACKED):
        print((ql, sl, ns), algo, check("rand", algo, qq, ss))
    if ql <= 2048:
        print((ql, sl, ns), "nonpacked", check("rand", B.BITPAL_NONPACKED, qq, ss))
    for sch in [(1, -1, -1), (1, -3, -2)]:
        print((ql, sl, ns), sch, check("rand", B.BITPAL_PACKED, qq, ss, match=sch[0], mismatch=sch[1], gap=sch[2]))
for L, e in [(100, 5), (100, 15), (100, 16), (100, 31), (64, 3), (50, 5), (250, 7), (640, 20), (333, 31), (1000, 10)]:
    qq = R.random_rows(rng, 2, L, with_n=0.005)
    ss = np.concatenate([R.mutate_rows(rng, qq[0, :L], 300, 2 * e + 2), R.indel_rows(rng, qq[1, :L], 200, e + 3), R.random_rows(rng, 77, L, with_n=0.01)])
    inb = ((L - 1) // 64 + 1) < ((L - e + 63) // 64 + 1)
    print("banded", (L, e), "ref-in-bounds", inb, check("banded", B.BANDED_MYERS, qq, ss, threshold=e))
print("BAD =", bad)

# rough timings, device resident
def time_config(name, algo, ql, sl, ns, reps=5, **kw):
    p = B.Params.default(algo, **kw)
    qq = R.random_rows(rng, 1, ql); ss = R.random_rows(rng, ns, sl)
    d_rows = torch.from_numpy(ss.reshape(-1)).cuda()
    d_packed = torch.empty(B.packed_bytes(sl, ns), dtype=torch.uint8, device="cuda")
    d_res = torch.empty(ns * 2, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for r in range(reps):
        ev[0].record(); B.pack_subjects_device(p, d_rows.data_ptr(), sl, ns, d_packed.data_ptr(), 0, st); ev[1].record()
        B.align_device(p, qq, d_packed.data_ptr(), sl, ns, d_res.data_ptr(), ns, 0, st); ev[2].record()
        torch.cuda.synchronize()
    tp, ta = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    cells = ql * sl * ns
    print(f"{name}: {B.kernel_name(p, ql, sl)} pack {tp:.3f} ms align {ta:.3f} ms -> {cells / ta / 1e6:.1f} GCUPS (kernel), {cells/(ta+tp)/1e6:.1f} with pack")
    t0 = time.time(); out = B.align_batch(p, qq, ss); t1 = time.time()
    print(f"   host-buffer call {1e3*(t1-t0):.2f} ms -> {cells/(t1-t0)/1e9:.1f} GCUPS e2e (pageable)")

ops, mhz = B.int_peak(0)
print(f"int peak {ops/1e12:.2f} T lane-op/s at {mhz:.0f} MHz")
time_config("C2", B.BITPAL_PACKED, 150, 150, 1_000_000)
time_config("C2np", B.BITPAL_NONPACKED, 150, 150, 200_000)
time_config("myers150", B.MYERS_GLOBAL, 150, 150, 1_000_000)
time_config("C1big", B.MYERS_GLOBAL, 500, 500, 200_000)
time_config("C3", B.BANDED_MYERS, 100, 100, 2_000_000, threshold=5)
time_config("C4", B.MYERS_SEMIGLOBAL, 1000, 1000, 100_000)
time_config("C5", B.BITPAL_PACKED, 5000, 5000, 4096)
ops, mhz = B.int_peak(0)
print(f"int peak {ops/1e12:.2f} T lane-op/s at {mhz:.0f} MHz")
