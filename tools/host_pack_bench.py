import sys, time, numpy as np, ctypes as C
from pathlib import Path; R = Path(__file__).resolve().parent.parent; sys.path[:0] = [str(R), str(R / 'tools'), str(R / 'tests')]
import bgsa_b200 as B, synth
lib = B.load()
for cfg,n,algo in (("C2",1_000_000,3),("C3",2_000_000,2),("C4",100_000,1)):
    q,s = synth.make(cfg,n)
    p = B.Params.default(algo, threshold=5)
    out = np.zeros(B.packed_bytes(s.shape[1]-1, n), dtype=np.uint8)
    best=1e9
    for r in range(5):
        t=time.perf_counter(); lib.bgsa_pack_subjects_host(C.byref(p), s.ctypes.data, s.shape[1]-1, n, out.ctypes.data); best=min(best,time.perf_counter()-t)
    print(cfg, B.host_pack_info(), f"{s.nbytes/best/1e9:.1f} GB/s in, {best*1e3:.1f} ms")
