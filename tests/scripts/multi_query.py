"""Many-queries mode (reference: up to REF_BUCKET_COUNT = 100 queries per bucket, cal_cpu.c:210-216): throughput of
nq queries x ns subjects through the host-buffer API and the device-resident API."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bgsa_b200 as B, refutil as R

rng = np.random.default_rng(1)
for algo, name, ql, sl, nq, ns in [(B.BITPAL_PACKED, "bitpal", 150, 150, 100, 100_000), (B.MYERS_GLOBAL, "myers", 150, 150, 100, 100_000),
                                   (B.BITPAL_PACKED, "bitpal", 150, 150, 3, 1_000_000), (B.MYERS_GLOBAL, "myers", 500, 500, 100, 20_000),
                                   (B.BANDED_MYERS, "banded", 100, 100, 100, 200_000)]:
    kw = {"threshold": 5} if algo == B.BANDED_MYERS else {}
    p = B.Params.default(algo, **kw)
    q = R.random_rows(rng, nq, ql); s = R.random_rows(rng, ns, sl)
    esz = 1 if algo == B.BANDED_MYERS else 2
    d_rows = torch.from_numpy(s.reshape(-1)).cuda()
    d_packed = torch.empty(B.packed_bytes(sl, ns), dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(nq * ns * esz, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    B.pack_subjects_device(p, d_rows.data_ptr(), sl, ns, d_packed.data_ptr(), 0, st)
    ts = []
    for r in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); B.align_device(p, q, d_packed.data_ptr(), sl, ns, d_res.data_ptr(), ns, 0, st); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts[1:]))
    cells = ql * sl * nq * ns
    h = torch.from_numpy(s.reshape(-1)).pin_memory().numpy().reshape(s.shape)
    out = torch.empty(nq * ns * esz, dtype=torch.uint8).pin_memory().numpy().view(np.int8 if esz == 1 else np.int16).reshape(nq, ns)
    for _ in range(2): B.align_batch(p, q, h, out=out)
    t0 = time.perf_counter()
    for _ in range(3): B.align_batch(p, q, h, out=out)
    te = (time.perf_counter() - t0) / 3
    chk = (out[:2, :2000] == R.oracle_batch({B.BITPAL_PACKED: 3, B.MYERS_GLOBAL: 0, B.BANDED_MYERS: 2}[algo], q[:2], s[:2000], e=5)).all()
    print(f"{name:7s} {nq:4d} q x {ns:8d} s x {ql} bp: kernel {t:8.3f} ms = {cells/t/1e6:9.0f} GCUPS; e2e {te*1e3:8.3f} ms = {cells/te/1e9:9.0f} GCUPS; sample parity {bool(chk)}", flush=True)
