"""Kernel-only timings (device-resident inputs) of the library named by BGSA_B200_LIB on a fixed set of
workloads; one line per workload.  Used for A/B comparisons of kernel variants on the GPU box:
    BGSA_B200_LIB=bgsa_b200/libX.so python tools/ab_kernels.py [tag]
Each workload is also checked against a CRC of the scores so that two variants can be compared for
bit-equality across processes."""
import sys, zlib
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import numpy as np, torch
import bgsa_b200 as B, synth

tag = sys.argv[1] if len(sys.argv) > 1 else "lib"
only = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
rng = np.random.default_rng(99)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def rows(n, ln):
    r = np.empty((n, ln + 1), dtype=np.uint8)
    r[:, :ln] = ACGT[rng.integers(0, 4, size=(n, ln), dtype=np.uint8)]
    r[:, ln] = 10
    return r


def run(name, algo, ql, sl, ns, reps=7, **kw):
    if only and name not in only:
        return
    p = B.Params.default(algo, **kw)
    if name in synth.CONFIGS and synth.CONFIGS[name]["qlen"] == ql:
        qq, ss = synth.make(name, ns)
    else:
        qq, ss = rows(1, ql), rows(ns, sl)
    esz = 1 if algo == B.BANDED_MYERS else 2
    d_rows = torch.from_numpy(ss.reshape(-1)).cuda()
    d_packed = torch.empty(B.packed_bytes(sl, ns), dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(ns * esz, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    tp, ta = [], []
    for r in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(); B.pack_subjects_device(p, d_rows.data_ptr(), sl, ns, d_packed.data_ptr(), 0, st); ev[1].record()
        B.align_device(p, qq, d_packed.data_ptr(), sl, ns, d_res.data_ptr(), ns, 0, st); ev[2].record()
        torch.cuda.synchronize()
        if r >= 2:
            tp.append(ev[0].elapsed_time(ev[1])); ta.append(ev[1].elapsed_time(ev[2]))
    tf = []
    if B.rows_kernel_name(p, ql, sl)[1]:       # the one-kernel path (ASCII rows in, one launch)
        for r in range(reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record(); B.align_rows_device(p, qq, d_rows.data_ptr(), sl, ns, d_res.data_ptr(), ns, 0, st); ev[1].record()
            torch.cuda.synchronize()
            if r >= 2:
                tf.append(ev[0].elapsed_time(ev[1]))
    crcf = zlib.crc32(d_res.cpu().numpy().tobytes())
    if tf:      # leave the two-kernel path's scores in d_res for the crc column
        B.pack_subjects_device(p, d_rows.data_ptr(), sl, ns, d_packed.data_ptr(), 0, st)
        B.align_device(p, qq, d_packed.data_ptr(), sl, ns, d_res.data_ptr(), ns, 0, st)
        torch.cuda.synchronize()
    crc = zlib.crc32(d_res.cpu().numpy().tobytes())
    cells = ql * sl * ns
    a, pk = float(np.median(ta)), float(np.median(tp))
    print(f"{tag} {name:10s} {B.kernel_name(p, ql, sl):52s} pack {pk:8.3f} ms  align {a:9.3f} ms  "
          f"{cells / a / 1e6:10.1f} GCUPS  crc {crc:08x}" + (f"  rows-kernel {float(np.median(tf)):8.3f} ms {cells / float(np.median(tf)) / 1e6:10.1f} GCUPS crc {crcf:08x}" if tf else ""), flush=True)


ops, mhz = B.int_peak(0)
print(f"{tag} int peak {ops/1e12:.2f} T lane-op/s at {mhz:.0f} MHz")
run("C2", B.BITPAL_PACKED, 150, 150, 1_000_000)
run("myers150", B.MYERS_GLOBAL, 150, 150, 1_000_000)
run("myers64", B.MYERS_GLOBAL, 64, 64, 2_000_000)
run("C1big", B.MYERS_GLOBAL, 500, 500, 300_000)
run("C3", B.BANDED_MYERS, 100, 100, 10_000_000, threshold=5)
run("C3s", B.BANDED_MYERS, 100, 100, 10_000_000, threshold=5)
run("C4", B.MYERS_SEMIGLOBAL, 1000, 1000, 300_000)
run("C5", B.BITPAL_PACKED, 5000, 5000, 8192, reps=4)
run("C2np", B.BITPAL_NONPACKED, 150, 150, 300_000)
run("bp111", B.BITPAL_PACKED, 150, 150, 1_000_000, match=1, mismatch=-1, gap=-1)
run("C2semi", B.BITPAL_PACKED_SEMIGLOBAL, 150, 150, 1_000_000)
run("C5semi", B.BITPAL_PACKED_SEMIGLOBAL, 5000, 5000, 8192, reps=4)
run("myers5k", B.MYERS_GLOBAL, 5000, 5000, 65536, reps=4)
run("myers2k", B.MYERS_GLOBAL, 2000, 2000, 262144, reps=4)
run("myers20k", B.MYERS_GLOBAL, 20000, 1000, 131072, reps=4)
run("C4semi2k", B.MYERS_SEMIGLOBAL, 2000, 1000, 131072, reps=4)
