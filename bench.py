#!/usr/bin/env python
"""bench.py -- GCUPS of the BGSA one-query-vs-many-subjects hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C3|C4|C5|myers150] [--impl reference]

A "step" is one pass of the hot path (pack + align kernels; banded Myers: one fused kernel) over one batch of synthetic subjects
(tools/synth.py, BASELINE.json configs).  Default workload = C2 = configs[1]: BitPAl packed
2/-3/-5 global, 1 query x 1M subjects x 150 bp (the configuration the metric is quoted on).

One JSON line on stdout (rank 0):
  value      whole-job GCUPS with the ASCII subject rows already resident in HBM (pack + align kernels),
             CUDA events, max over ranks; N > 1: every rank owns its own shard of equal size (weak).
  e2e        the same metric through the reference-facing C-ABI call bgsa_align_batch with PINNED HOST
             buffers: H2D of the rows and D2H of the scores are inside the timed region every step.
  roofline   the align kernel against the INT32 ALU-pipe roofline: achieved = algorithmic lane-ops
             (SURVEY.md section 8d instruction model) / its measured duration; peak = LOP3 issue rate
             measured live by bgsa_int_peak (MEASURED_PEAKS.json has no integer figure).
             roofline.pipe_utilisation_ncu / roofline.traffic come from the committed ncu capture of the same kernel.
  cpu_baseline  the unmodified reference (oracle/_ref, built from the reference sources) timed on this
             box's host cores on the same workload (a bounded, strided sample of it when it is larger than 1M subjects).
  parity     mismatches between the scores the timed e2e steps produced and the reference's scores for that sample.
`--impl reference` prints the reference arm: the reference's own CPU code for the path
(<arch>_handle_reads + <arch>_cal_align_score), all host threads, same config/metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import zlib
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (ROOT, ROOT / "tests", ROOT / "tools"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

import numpy as np  # noqa: E402

import synth  # noqa: E402

WORKLOADS = {
    # name: synth config, algorithm (bgsa_algo_t), per-rank subject count, reference variant, params
    "C2": dict(cfg="C2", algo=3, count=1_000_000, ref="bitpal_avx512", ref_alt="bitpal_avx2", kw={},
               desc="BitPAl packed M=2 I=-3 G=-5 global, 1 query x 1M synthetic 150bp subjects"),
    "C3": dict(cfg="C3", algo=2, count=10_000_000, ref="banded_cpu", ref_alt=None, kw={"threshold": 5},
               desc="banded Myers verification e=5, 1 query x 10M synthetic 100bp subjects"),
    "C4": dict(cfg="C4", algo=1, count=1_000_000, ref="semiglobal_cpu", ref_alt=None, kw={},
               desc="semi-global Myers, 1 query x 1M synthetic 1000bp subjects"),
    "C5": dict(cfg="C5", algo=3, count=125_000, ref="bitpal_avx512", ref_alt="bitpal_avx2", kw={},
               desc="BitPAl global 5kbp query x 5kbp subjects, 125k subjects per GPU (1M over 8 GPUs)"),
    "myers150": dict(cfg="C2", algo=0, count=1_000_000, ref="myers_sse", ref_alt="myers_cpu", kw={},
                     desc="Myers unit-cost global, 1 query x 1M synthetic 150bp subjects"),
}
ALGO_NAME = {0: "myers_global", 1: "myers_semiglobal", 2: "banded_myers", 3: "bitpal_packed", 4: "bitpal_nonpacked"}


def alg_ops_per_cell(algo: int, qlen: int, slen: int) -> float:
    """algorithmic ALU-pipe instructions per DP cell (the roofline model)."""
    W = (qlen + 31) // 32
    if algo in (0, 1):
        return (10 * W + 5) / qlen
    if algo in (3, 4):
        return (80 * W + 5) / qlen
    return 16.0 / slen          # banded: 16 per band row, rows = query length, per NOMINAL cell


def alg_bytes_per_subject(algo: int, slen: int) -> float:
    return slen / 4.0 + (1 if algo == 2 else 2)     # 2-bit bases in, one score out


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
def reference_run(wl: dict, query, subjects, min_seconds: float, max_runs: int):
    """Times the unmodified reference (oracle/_ref) on host cores: Peq build + kernel."""
    import refutil as R
    variant = wl["ref"] if R.ref_available(wl["ref"]) else wl["ref_alt"]
    if variant is None or not R.ref_available(variant):
        return None
    ref = R.reflib(variant)
    st = ref.prepare(query, subjects, e=wl["kw"].get("threshold", 5))
    t_handle, t_cal, runs = [], [], 0
    t_begin = time.perf_counter()
    while runs < max_runs and (runs < 2 or time.perf_counter() - t_begin < min_seconds):
        t0 = time.perf_counter(); ref.handle_reads(st); t1 = time.perf_counter(); scores = ref.cal_align_score(st); t2 = time.perf_counter()
        t_handle.append(t1 - t0); t_cal.append(t2 - t1); runs += 1
    cells = float(query.shape[1] - 1) * (subjects.shape[1] - 1) * subjects.shape[0] * query.shape[0]
    th, tc = min(t_handle), min(t_cal)
    kind = "port" if variant == "semiglobal_cpu" else "reference"
    return dict(variant=variant, kind=kind, cores=ref.threads, runs=runs, cells=cells, t_handle=th, t_cal=tc,
                gcups_path=cells / (th + tc) / 1e9, gcups_cal=cells / tc / 1e9, scores=scores)


def _claim_stdout():
    """The contract is ONE JSON line on stdout: anything a library prints there (NCCL's version banner under
    NCCL_DEBUG=VERSION, for one) goes to stderr instead; the returned writer emits on the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real, (json.dumps(obj) + "\n").encode())
    return emit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--count", type=int, default=None, help="subjects per rank (default: the config's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    wl = WORKLOADS[args.workload]
    count = args.count or wl["count"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and args.impl == "ours":
        # convenience: re-launch ourselves one rank per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", "29517", __file__] + sys.argv[1:])
    emit = _claim_stdout()

    cfg = synth.CONFIGS[wl["cfg"]]
    qlen, slen = cfg["qlen"], cfg["slen"]
    config = {"workload": f"{args.workload}: {wl['desc']}", "algorithm": ALGO_NAME[wl["algo"]], "query_len": qlen,
              "subject_len": slen, "subjects_per_gpu": count, "n_queries": 1,
              "l2": "inputs larger than L2: %.0f MB of ASCII rows per step vs 126 MB" % (count * (slen + 1) / 1e6)}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        query, subjects = synth.make(wl["cfg"], count)
        res = reference_run(wl, query, subjects, min_seconds=20.0, max_runs=max(args.steps + args.warmup, 3))
        if res is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref not built for this host CPU"})
            return
        v = res["gcups_path"]
        emit({
            "impl": "reference", "metric": "GCUPS", "value": v, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * (res["t_handle"] + res["t_cal"]), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": res["cores"], "kind": res["kind"], "variant": res["variant"],
                             "sample": f"full workload ({subjects.shape[0]} subjects), best of {res['runs']} runs, Peq build + kernel",
                             "cal_only_gcups": res["gcups_cal"]},
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ---------------------------------------------------------------- our arm
    import torch
    import bgsa_b200 as B
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B.load()
    # one process per GPU: allocate this rank's pinned buffers on the GPU's own NUMA node (BGSA_NO_NUMA_BIND=1 to skip)
    numa_node = -1 if os.environ.get("BGSA_NO_NUMA_BIND") else B.bind_thread_to_device(local_rank)
    config["host_numa_node"] = numa_node
    params = B.Params.default(wl["algo"], **wl["kw"])
    # every rank owns its own contiguous shard of equal size (weak scaling; no data-path collective)
    # (rank r > 0 draws its own subjects from the same recipe -- same query, same mix -- so that every GPU does the
    #  same kind of work as the single GPU of the N = 1 run)
    query, subjects = synth.make(wl["cfg"], count, shard=rank)
    ns = subjects.shape[0]
    cells = float(qlen) * slen * ns
    esize = 1 if wl["algo"] == B.BANDED_MYERS else 2
    stream = torch.cuda.current_stream().cuda_stream
    dev = local_rank

    int_peak, sm_mhz_probe = B.int_peak(dev)

    # resident buffers
    h_rows = torch.from_numpy(subjects.reshape(-1)).pin_memory()
    d_rows = h_rows.cuda(non_blocking=True)
    d_packed = torch.empty(B.packed_bytes(slen, ns), dtype=torch.uint8, device="cuda")
    d_res = torch.empty(ns * esize, dtype=torch.uint8, device="cuda")
    h_res = torch.empty(ns * esize, dtype=torch.uint8).pin_memory()
    subj_pinned = h_rows.numpy().reshape(ns, slen + 1)
    out_pinned = h_res.numpy().view(np.int8 if esize == 1 else np.int16).reshape(1, ns)

    # resident step: ASCII rows in HBM -> scores.  Pack + align for the transposed-DP algorithms (events around the align
    # kernel give its own duration); banded Myers is ONE fused kernel (bgsa_align_rows_device), timed as a whole.
    fused = wl["algo"] == B.BANDED_MYERS

    def step_resident(ev=None):
        if fused:
            if ev:
                ev[0].record()
            B.align_rows_device(params, query, d_rows.data_ptr(), slen, ns, d_res.data_ptr(), ns, dev, stream)
        else:
            B.pack_subjects_device(params, d_rows.data_ptr(), slen, ns, d_packed.data_ptr(), dev, stream)
            if ev:
                ev[0].record()
            B.align_device(params, query, d_packed.data_ptr(), slen, ns, d_res.data_ptr(), ns, dev, stream)
        if ev:
            ev[1].record()

    def step_e2e():
        B.align_batch(params, query, subj_pinned, device=dev, out=out_pinned)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident timing
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = B.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        step_resident(evs[i])
    t_end.record()
    barrier()
    launches = B.launch_count() - launches0
    ms_total = t_start.elapsed_time(t_end)
    align_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    # ---- end-to-end timing (host pinned buffers, H2D + D2H inside)
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    int_peak2, _ = B.int_peak(dev)
    int_peak = max(int_peak, int_peak2)

    if dist is not None:
        t = torch.tensor([ms_total, e2e_s, align_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s, align_ms = (float(x) for x in t.cpu())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    value = cells * world / (ms_per_step * 1e-3) / 1e9
    e2e_value = cells * world / (e2e_s / args.steps) / 1e9
    ops_cell = alg_ops_per_cell(wl["algo"], qlen, slen)
    achieved = ops_cell * cells / (align_ms * 1e-3)               # lane-ops/s of the align kernel, per GPU
    hbm_bytes = alg_bytes_per_subject(wl["algo"], slen) * ns
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # roofline.traffic: DRAM bytes of the align kernel per launch, from the committed ncu --set full capture of the
    # same kernel (profiles/ncu_summary.json, written by tools/ncu_summarize.py), scaled to this launch's subjects
    traffic, pipe_ncu = None, None
    try:
        prof = json.loads((ROOT / "profiles" / "ncu_summary.json").read_text()).get(args.workload, {})
        if "dram_bytes_per_subject" in prof:
            traffic = prof["dram_bytes_per_subject"] * ns
        if "alu_pipe_pct" in prof:
            # frac is measured against the survey's instruction MODEL; a kernel that needs fewer instructions per cell
            # than the model reads above 1.0 while the pipe itself cannot exceed 100 % -- this is the pipe's own counter
            pipe_ncu = {"alu_pipe_busy": prof["alu_pipe_pct"] / 100.0, "issue_slots_busy": prof["issue_active_pct"] / 100.0,
                        "fma_pipe_busy": prof["fma_pipe_pct"] / 100.0, "kernel": prof.get("kernel"), "source": prof.get("source")}
    except (OSError, ValueError):
        pass
    line = {
        "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": config,
        "kernel": B.kernel_name(params, qlen, slen) + (" (fused: ASCII tile -> shared-memory strip -> band)" if fused else ""),
        "value_align_kernel_only": cells * world / (align_ms * 1e-3) / 1e9,
        "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": int(ns * (slen + 1)), "d2h_bytes_per_step": int(ns * esize),
                "ms_per_step": 1e3 * e2e_s / args.steps, "api": "bgsa_align_batch (pinned host rows in, host scores out)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "int_alu", "achieved": achieved / 1e12, "peak": int_peak / 1e12, "unit": "Tlane-op/s",
                     "frac": achieved / int_peak, "traffic": traffic, "pipe_utilisation_ncu": pipe_ncu,
                     "kernel_ms": align_ms, "ops_per_cell_model": ops_cell,
                     "peak_source": "measured live: LOP3 issue-rate probe bgsa_int_peak (MEASURED_PEAKS.json has no integer peak); "
                                    "nominal 148 SM x 64 lanes x 1.965 GHz = 18.6",
                     "hbm": {"achieved": hbm_bytes / (align_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": hbm_bytes / (align_ms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650"}},
        "clocks": clocks,
        "sm_mhz_probe": sm_mhz_probe,
    }
    if world == 1 and not args.no_cpu_baseline:
        # bounded sample for the CPU: every (ns / 1M)-th subject, so that it spans the whole workload (C3: both the
        # near-identical and the random half)
        sample_idx = np.arange(ns) if ns <= 1_000_000 else np.arange(0, ns, ns // 1_000_000)[:1_000_000]
        res = reference_run(wl, query, subjects if ns <= 1_000_000 else np.ascontiguousarray(subjects[sample_idx]),
                            min_seconds=10.0, max_runs=10)
        if res is not None:
            line["cpu_baseline"] = {"value": res["gcups_path"], "unit": "GCUPS", "cores": res["cores"], "kind": res["kind"],
                                    "variant": res["variant"], "cal_only_gcups": res["gcups_cal"],
                                    "sample": f"{len(sample_idx)} subjects of the same workload"
                                              f"{'' if ns <= 1_000_000 else ' (every %d-th)' % (ns // 1_000_000)}, best of {res['runs']} runs, "
                                              f"Peq build ({res['t_handle']:.3f} s) + kernel ({res['t_cal']:.3f} s)"}
            # the checker at work: the scores the timed e2e steps left in the pinned result buffer against the
            # reference's scores for the same subjects (bit-exact or the run is worthless)
            nref = res["scores"].shape[1]
            mism = int((out_pinned[:, sample_idx[:nref]] != res["scores"]).sum())
            line["parity"] = {"against": res["variant"], "subjects_compared": int(nref), "mismatches": mism,
                              "crc32_all_scores": "%08x" % zlib.crc32(out_pinned.tobytes())}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
