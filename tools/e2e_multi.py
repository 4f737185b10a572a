"""End-to-end batch entry with SEVERAL ranks sharing one host (torchrun, one rank per GPU): bgsa_align_batch on pinned rows,
front-end modes BGSA_HOST_PACK = 0 (ASCII over the link, device pack) / auto, all ranks timed together between barriers.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/e2e_multi.py [workloads] [modes]
Prints, per workload and mode, the slowest rank's ms per batch and the aggregate input rate."""
import os, sys, time
from pathlib import Path
R = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(R), str(R / "tools"), str(R / "tests")]
import numpy as np, torch, torch.distributed as dist
import bgsa_b200 as B, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
W = {"C2": ("C2", 3, 1_000_000, {}), "C3": ("C3", 2, 10_000_000, {"threshold": 5}), "myers150": ("C2", 0, 1_000_000, {}),
     "C4": ("C4", 1, 300_000, {})}
only = sys.argv[1].split(",") if len(sys.argv) > 1 else ["C2", "C3", "myers150"]
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "auto"]      # also: 1, 2, model (= auto without the tuner)
B.load()
B.bind_thread_to_device(local)
if rank == 0:
    print(f"ranks {world}, host cores {os.cpu_count()}, pool per rank {B.host_pack_info()}", flush=True)
for name in only:
    cfg, algo, n, kw = W[name]
    q, s = synth.make(cfg, n, shard=rank)
    p = B.Params.default(algo, **kw)
    h = torch.from_numpy(s.reshape(-1)).pin_memory()
    sp = h.numpy().reshape(s.shape)
    es = 1 if algo == B.BANDED_MYERS else 2
    hr = torch.empty(n * es, dtype=torch.uint8).pin_memory()
    out = hr.numpy().view(np.int8 if es == 1 else np.int16).reshape(1, n)
    for mode in modes:
        os.environ.pop("BGSA_HOST_PACK", None)
        os.environ.pop("BGSA_HOST_PACK_NO_TUNING", None)
        if mode == "model":
            os.environ["BGSA_HOST_PACK_NO_TUNING"] = "1"
        elif mode != "auto":
            os.environ["BGSA_HOST_PACK"] = mode
        for _ in range(3):
            B.align_batch(p, q, sp, device=local, out=out)
        if world > 1:
            dist.barrier()
        reps = 24
        shares = []
        t = time.perf_counter()
        for _ in range(reps):
            B.align_batch(p, q, sp, device=local, out=out)
            shares.append(B.batch_front_end(local, 0))
        dt = (time.perf_counter() - t) / reps
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt[0])
        if rank == 0:
            cells = (q.shape[1] - 1) * (s.shape[1] - 1) * n * world
            print(f"{name:9s} N={world} host_pack={mode:4s} {dt * 1e3:8.3f} ms  {cells / dt / 1e9:9.0f} GCUPS  rows {s.nbytes * world / dt / 1e9:6.1f} GB/s aggregate  rank 0 host-pack share per job: {' '.join('%.2f' % v for v in shares)}", flush=True)
    del h, hr
if world > 1:
    dist.destroy_process_group()
