#!/usr/bin/env python
"""bench.py -- GCUPS of the BGSA one-query-vs-many-subjects hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--only-headline] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic subjects (tools/synth.py, BASELINE.json configs).
The HEADLINE workload (top-level keys of the JSON line) is C2 = configs[1]: BitPAl packed 2/-3/-5 global,
1 query x 1M subjects x 150 bp -- the configuration the metric is quoted on.  The same line then carries, under
"workloads", every other configuration of BASELINE.json measured in the same run by the same code:
    C3 (banded e=5, 10M x 100 bp), C3s (the same subjects in random order), C4 (semi-global Myers, 1M x 1000 bp),
    C5 (BitPAl 5 kbp x 5 kbp, 125k subjects per GPU = 1M over 8 GPUs), myers150 (Myers global 150 bp) and
    C2np (BitPAl NON-packed 150 bp),
each with value / kernel time / roofline / e2e / cpu_baseline / parity.  `--only-headline` skips them,
`--workload X` makes X the headline.

Per workload:
  value      whole-job GCUPS with the ASCII subject rows already resident in HBM (bgsa_align_rows_device: ONE kernel fed
             with the ASCII rows for short reads -- align_rows_kernel, fused banded kernel -- else pack + align kernels),
             CUDA events, max over ranks; N > 1: every rank owns its own shard of equal size (weak).  Where the one-kernel
             path runs, "two_kernel_path" reports pack + align on packed tiles beside it.
  e2e        the same metric through the reference-facing C-ABI call bgsa_align_batch with PINNED HOST buffers:
             H2D of the rows and D2H of the scores are inside the timed region every step.  The library's front end may
             let the host threads 2-bit-pack a share of the chunks instead of shipping them as ASCII (tuned job by job on
             measured throughput; the share of the last step is reported); h2d_bytes_per_step counts the ASCII size.
  roofline   the align kernel against the INT32 ALU-pipe roofline.  frac = SURVEY.md section 8d's MODEL instruction
             count / measured duration / peak (can exceed 1 where the kernel needs fewer instructions than the model);
             frac_sass = the same with the ALU-pipe instructions the kernel REALLY executes per cell, counted from
             its SASS hot loop by tools/sass_budget.py (profiles/sass_budget.json) -- an estimate of the pipe's
             utilisation, <= 1 by construction.  peak = LOP3 issue rate measured live by bgsa_int_peak
             (MEASURED_PEAKS.json has no integer figure).  traffic / pipe_utilisation_ncu: from the committed ncu
             capture of the same kernel (profiles/ncu_summary.json; a profiler cannot run inside a timed bench).
  cpu_baseline  the unmodified reference (oracle/_ref, built from the reference sources) timed on this box's host
             cores on a bounded sample of the same workload (rank 0, N = 1 only).
  parity     EVERY rank compares the scores its timed e2e steps produced with the reference's scores for a sample of
             ITS OWN shard; mismatches are summed over ranks.  Bit-exact or the run is worthless.
At N > 1 rank 0 finally drives all N devices from ONE process the way the product does (aligner -g N, and the
reference's cal_mic.c:459-481): contiguous subject ranges (bgsa_b200/sharding.py), bgsa_align_batch_submit per
device, one pinned result buffer, device-major -- on the C5 set ("single_process").
`--impl reference` prints the reference arm: the reference's own CPU code for the path
(<arch>_handle_reads + <arch>_cal_align_score), all host threads, same config/metric.
Timing statistics: our arm reports the MEAN over the K timed steps; the reference arm the BEST of its runs.
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
    # name: synth config, algorithm (bgsa_algo_t), per-rank subject count, reference variant, params,
    #       ref_sample = subjects the CPU reference is run on (timing at N = 1, parity on every rank), sass = kernel
    #       name in profiles/sass_budget.json, max_steps = cap on the timed steps when the workload is not the headline
    "C2": dict(cfg="C2", algo=3, count=1_000_000, ref="bitpal_avx512", ref_alt="bitpal_avx2", kw={}, ref_sample=1_000_000,
               sass="C2_rows_bitpal_packed_K5", desc="BitPAl packed M=2 I=-3 G=-5 global, 1 query x 1M synthetic 150bp subjects"),
    "C3": dict(cfg="C3", algo=2, count=10_000_000, ref="banded_cpu", ref_alt=None, kw={"threshold": 5}, ref_sample=1_000_000,
               sass="C3_banded_fused", desc="banded Myers verification e=5, 1 query x 10M synthetic 100bp subjects"),
    "C3s": dict(cfg="C3s", algo=2, count=10_000_000, ref="banded_cpu", ref_alt=None, kw={"threshold": 5}, ref_sample=1_000_000,
                sass="C3_banded_fused", desc="C3 with its near-identical and random halves interleaved at random (order independence)"),
    "C4": dict(cfg="C4", algo=1, count=1_000_000, ref="semiglobal_cpu", ref_alt=None, kw={}, ref_sample=200_000,
               sass="C4_myers_semi_K32", desc="semi-global Myers, 1 query x 1M synthetic 1000bp subjects"),
    "C5": dict(cfg="C5", algo=3, count=125_000, ref="bitpal_avx512", ref_alt="bitpal_avx2", kw={}, ref_sample=16_000, max_steps=5,
               sass="C5_bitpal_packed_K10_L16", desc="BitPAl global 5kbp query x 5kbp subjects, 125k subjects per GPU (1M over 8 GPUs)"),
    "myers150": dict(cfg="C2", algo=0, count=1_000_000, ref="myers_sse", ref_alt="myers_cpu", kw={}, ref_sample=1_000_000,
                     sass="myers150_rows_K5", desc="Myers unit-cost global, 1 query x 1M synthetic 150bp subjects"),
    "C2np": dict(cfg="C2", algo=4, count=1_000_000, ref="bitpal_avx512", ref_alt="bitpal_avx2", kw={}, ref_sample=1_000_000,
                 sass="C2np_rows_K5", desc="BitPAl NON-packed M=2 I=-3 G=-5 global, 1 query x 1M synthetic 150bp subjects "
                                                   "(reference: the packed AVX-512 build -- same scores by definition)"),
}
EXTRA_ORDER = ["C3", "C3s", "C4", "C5", "myers150", "C2np"]
ALGO_NAME = {0: "myers_global", 1: "myers_semiglobal", 2: "banded_myers", 3: "bitpal_packed", 4: "bitpal_nonpacked"}


def alg_ops_per_cell(algo: int, qlen: int, slen: int) -> float:
    """algorithmic ALU-pipe instructions per DP cell (the roofline model)."""
    W = (qlen + 31) // 32
    if algo in (0, 1):
        return (10 * W + 5) / qlen
    if algo in (3, 4):
        return (80 * W + 5) / qlen
    return 16.0 / slen          # banded: 16 per band row, rows = query length, per NOMINAL cell


def alg_bytes_per_subject(algo: int, slen: int, ascii_in: bool) -> float:
    """HBM bytes the align kernel must move per subject: the one-kernel paths read the ASCII rows (one byte per base plus
    the row end), the kernels on packed tiles 2 bits per base; one score out."""
    return (slen + 1.0 if ascii_in else slen / 4.0) + (1 if algo == 2 else 2)


def _git_sha() -> str | None:
    try:
        return subprocess.run(["git", "-C", str(ROOT), "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                              timeout=5).stdout.strip() or None
    except (OSError, subprocess.SubprocessError):
        return None


def _load_json(path: Path) -> dict:
    try:
        return json.loads(path.read_text())
    except (OSError, ValueError):
        return {}


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
        # under load = the samples in the upper half of the power range (the sampler also sees the idle gaps between
        # workloads while data is generated on the host)
        busy = [s for s, p in zip(sm, power) if power and p >= 0.5 * (min(power) + max(power))] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "samples_under_load": len(busy),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
def reference_run(wl: dict, query, subjects, min_seconds: float, max_runs: int, threads: int = 0):
    """Times the unmodified reference (oracle/_ref) on host cores: Peq build + kernel."""
    import refutil as R
    variant = wl["ref"] if R.ref_available(wl["ref"]) else wl["ref_alt"]
    if variant is None or not R.ref_available(variant):
        return None
    ref = R.reflib(variant) if not threads else R.RefLib(variant, threads=threads)
    st = ref.prepare(query, subjects, e=wl["kw"].get("threshold", 5))
    t_handle, t_cal, runs = [], [], 0
    t_begin = time.perf_counter()
    while runs < max_runs and (runs < (2 if max_runs > 1 else 1) or time.perf_counter() - t_begin < min_seconds):
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


def log(msg: str):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def workload_config(name: str, count: int) -> dict:
    wl = WORKLOADS[name]
    cfg = synth.CONFIGS[wl["cfg"]]
    return {"workload": f"{name}: {wl['desc']}", "algorithm": ALGO_NAME[wl["algo"]], "query_len": cfg["qlen"],
            "subject_len": cfg["slen"], "subjects_per_gpu": count, "n_queries": 1,
            "l2": "inputs larger than L2: %.0f MB of ASCII rows per step vs 126 MB" % (count * (cfg["slen"] + 1) / 1e6)}


def sample_indices(ns: int, want: int) -> np.ndarray:
    """A bounded sample that SPANS the shard (C3: both the near-identical and the random half)."""
    if ns <= want:
        return np.arange(ns)
    return np.arange(0, ns, ns // want)[:want]


# --------------------------------------------------------------------------------------------------
class Ctx:
    """Everything a workload measurement needs from the process: rank, device, torch, the library."""

    def __init__(self, args):
        import torch
        import bgsa_b200 as B
        self.torch, self.B, self.args = torch, B, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        self.dist, self.gloo = None, None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
            self.gloo = dist.new_group(backend="gloo")      # host-side waits (a NCCL barrier would spin on the GPUs)
        B.load()
        # one process per GPU: allocate this rank's pinned buffers on the GPU's own NUMA node (BGSA_NO_NUMA_BIND=1 to skip)
        self.numa_node = -1 if os.environ.get("BGSA_NO_NUMA_BIND") else B.bind_thread_to_device(self.local_rank)
        self.dev = self.local_rank
        self.stream = torch.cuda.current_stream().cuda_stream
        self.int_peak, self.sm_mhz_probe = B.int_peak(self.dev)
        self.peaks = _load_json(ROOT / "MEASURED_PEAKS.json")
        self.ncu = _load_json(ROOT / "profiles" / "ncu_summary.json")
        self.sass = _load_json(ROOT / "profiles" / "sass_budget.json")
        self.host_cores = os.cpu_count() or 1

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op="max"):
        if self.dist is None:
            return [float(v) for v in values]
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.cpu()]


def measure(ctx: Ctx, name: str, steps: int, warmup: int, count: int | None, cpu_seconds: float, data_cache: dict) -> dict:
    """One workload on every rank; returns the record (meaningful on rank 0; timing = max over ranks)."""
    torch, B = ctx.torch, ctx.B
    wl = WORKLOADS[name]
    cfg = synth.CONFIGS[wl["cfg"]]
    qlen, slen = cfg["qlen"], cfg["slen"]
    count = count or wl["count"]
    params = B.Params.default(wl["algo"], **wl["kw"])
    t_gen = time.perf_counter()
    # every rank owns its own contiguous shard of equal size (weak scaling; no data-path collective): rank r > 0
    # draws its own subjects from the same recipe -- same query, same mix
    key = (wl["cfg"], count, ctx.rank)
    if key in data_cache:
        query, subjects = data_cache[key]
    else:
        base = data_cache.get(("C3", count, ctx.rank)) if wl["cfg"] == "C3s" else None
        data_cache.clear()                      # one data set alive at a time
        if base is not None:                    # C3s = C3's subjects in random order: no need to draw them again
            query, subjects = base[0], synth.shuffled(base[1], "C3s")
            del base
        else:
            query, subjects = synth.make(wl["cfg"], count, shard=ctx.rank)
        data_cache[key] = (query, subjects)
    ns = subjects.shape[0]
    cells = float(qlen) * slen * ns
    esize = 1 if wl["algo"] == B.BANDED_MYERS else 2
    dev, stream = ctx.dev, ctx.stream
    if ctx.rank == 0:
        log(f"{name}: {ns} subjects x {slen} bp generated in {time.perf_counter() - t_gen:.1f} s")

    h_rows = torch.from_numpy(subjects.reshape(-1)).pin_memory()
    d_rows = h_rows.cuda(non_blocking=True)
    d_packed = torch.empty(B.packed_bytes(slen, ns), dtype=torch.uint8, device="cuda")
    d_res = torch.empty(ns * esize, dtype=torch.uint8, device="cuda")
    h_res = torch.empty(ns * esize, dtype=torch.uint8).pin_memory()
    subj_pinned = h_rows.numpy().reshape(ns, slen + 1)
    out_pinned = h_res.numpy().view(np.int8 if esize == 1 else np.int16).reshape(1, ns)

    # resident step: ASCII rows in HBM -> scores.  Pack + align for the transposed-DP algorithms (events around the align
    # kernel give its own duration); banded Myers is ONE fused kernel (bgsa_align_rows_device), timed as a whole.
    rows_kernel, fused = B.rows_kernel_name(params, qlen, slen)

    def step_resident(ev=None, one_kernel=fused):
        if one_kernel:
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

    # ---- resident timing
    for _ in range(warmup):
        step_resident()
    ctx.barrier()
    launches0 = B.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(steps):
        step_resident(evs[i])
    t_end.record()
    ctx.barrier()
    launches = B.launch_count() - launches0
    ms_total = t_start.elapsed_time(t_end)
    align_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    crc_resident = zlib.crc32(d_res.cpu().numpy().tobytes())
    # the two-kernel path (pack + align on packed tiles) beside the one-kernel path, where the latter is what runs
    two = None
    if fused and wl["algo"] != B.BANDED_MYERS:
        for _ in range(3):
            step_resident(one_kernel=False)
        evs2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t2s, t2e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t2s.record()
        for i in range(steps):
            step_resident(evs2[i], one_kernel=False)
        t2e.record()
        torch.cuda.synchronize()
        ms2 = t2s.elapsed_time(t2e) / steps
        a2 = float(np.mean([a.elapsed_time(b) for a, b in evs2]))
        two = {"kernels": "pack_stream_kernel + " + B.kernel_name(params, qlen, slen), "ms_per_step": ms2, "align_kernel_ms": a2,
               "value": cells / (ms2 * 1e-3) / 1e9, "value_align_kernel_only": cells / (a2 * 1e-3) / 1e9, "unit": "GCUPS (this rank)",
               "same_scores": zlib.crc32(d_res.cpu().numpy().tobytes()) == crc_resident}
    # ---- end-to-end timing (host pinned buffers, H2D + D2H inside)
    for _ in range(3):
        step_e2e()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    front_end_share = B.batch_front_end(dev, 0)
    crc_all = zlib.crc32(out_pinned.tobytes())

    # ---- the checker: the reference on a sample of THIS rank's shard (all ranks); timed on rank 0 at N = 1
    want = wl["ref_sample"] if ctx.world == 1 else min(wl["ref_sample"], max(2_000, wl["ref_sample"] // (4 * ctx.world)))
    sample_idx = sample_indices(ns, want)
    sample = subjects if len(sample_idx) == ns else np.ascontiguousarray(subjects[sample_idx])
    timing = ctx.world == 1 and not ctx.args.no_cpu_baseline
    threads = 0 if ctx.world == 1 else max(1, ctx.host_cores // ctx.world)
    res = None
    if not ctx.args.no_parity:
        res = reference_run(wl, query, sample, min_seconds=cpu_seconds if timing else 0.0, max_runs=10 if timing else 1, threads=threads)
    mism, compared = 0, 0
    if res is not None:
        nref = res["scores"].shape[1]
        mism = int((out_pinned[:, sample_idx[:nref]] != res["scores"]).sum())
        compared = int(nref)
    ms_total, e2e_s, align_ms = ctx.reduce([ms_total, e2e_s, align_ms], "max")
    mism_all, compared_all, checked_ranks = ctx.reduce([mism, compared, 1 if res is not None else 0], "sum")
    crcs = [crc_all]
    if ctx.dist is not None:
        t = torch.tensor([crc_all], dtype=torch.int64, device="cuda")
        gathered = [torch.zeros_like(t) for _ in range(ctx.world)]
        ctx.dist.all_gather(gathered, t)
        crcs = [int(g.item()) for g in gathered]

    world = ctx.world
    ms_per_step = ms_total / steps
    ops_cell = alg_ops_per_cell(wl["algo"], qlen, slen)
    achieved = ops_cell * cells / (align_ms * 1e-3)               # lane-ops/s of the align kernel, per GPU
    hbm_bytes = alg_bytes_per_subject(wl["algo"], slen, fused) * ns
    hbm_peak = ctx.peaks.get("hbm_gbs", 6650.0)
    prof = ctx.ncu.get(name if name != "C3s" else "C3", {})
    traffic, pipe_ncu = None, None
    if "dram_bytes_per_subject" in prof:
        traffic = prof["dram_bytes_per_subject"] * ns
    if "alu_pipe_pct" in prof:
        pipe_ncu = {"alu_pipe_busy": prof["alu_pipe_pct"] / 100.0, "issue_slots_busy": prof["issue_active_pct"] / 100.0,
                    "fma_pipe_busy": prof["fma_pipe_pct"] / 100.0, "kernel": prof.get("kernel"), "source": prof.get("source"),
                    "captured_at_git": prof.get("captured_at_git"),
                    "note": "committed ncu --set full capture of the same kernel (a profiler cannot run inside a timed bench)"}
    sass = ctx.sass.get(wl["sass"], {})
    ops_cell_sass = sass.get("ops_per_cell_sass")
    rec = {
        "config": workload_config(name, ns),
        "value": cells * world / (ms_per_step * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms_per_step, "steps": steps,
        "kernel": rows_kernel,
        "value_align_kernel_only": cells * world / (align_ms * 1e-3) / 1e9,
        "e2e": {"value": cells * world / (e2e_s / steps) / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(ns * (slen + 1)),
                "d2h_bytes_per_step": int(ns * esize), "ms_per_step": 1e3 * e2e_s / steps,
                "api": "bgsa_align_batch (pinned host rows in, host scores out)",
                "host_packed_share_of_chunks": front_end_share},
        "gpu_launches": int(launches),
        "roofline": {"bound": "int_alu", "achieved": achieved / 1e12, "peak": ctx.int_peak / 1e12, "unit": "Tlane-op/s",
                     "frac": achieved / ctx.int_peak, "traffic": traffic, "pipe_utilisation_ncu": pipe_ncu,
                     "kernel_ms": align_ms, "ops_per_cell_model": ops_cell, "ops_per_cell_sass": ops_cell_sass,
                     "frac_sass": (ops_cell_sass * cells / (align_ms * 1e-3) / ctx.int_peak) if ops_cell_sass else None,
                     "sass_source": "profiles/sass_budget.json (tools/sass_budget.py: ALU-pipe instructions of the hot loop per cell)"
                                    if ops_cell_sass else None,
                     "peak_source": "measured live: LOP3 issue-rate probe bgsa_int_peak (MEASURED_PEAKS.json has no integer peak); "
                                    "nominal 148 SM x 64 lanes x 1.965 GHz = 18.6",
                     "hbm": {"achieved": hbm_bytes / (align_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": hbm_bytes / (align_ms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in ctx.peaks else "fallback 6650"}},
        "parity": {"against": res["variant"] if res else None, "ranks_checked": int(checked_ranks), "subjects_compared": int(compared_all),
                   "mismatches": int(mism_all), "crc32_scores_per_rank": ["%08x" % (c & 0xffffffff) for c in crcs]},
    }
    if two is not None:
        rec["two_kernel_path"] = two
    if res is not None and timing:
        rec["cpu_baseline"] = {"value": res["gcups_path"], "unit": "GCUPS", "cores": res["cores"], "kind": res["kind"],
                               "variant": res["variant"], "cal_only_gcups": res["gcups_cal"],
                               "sample": f"{len(sample_idx)} subjects of the same workload"
                                         f"{'' if len(sample_idx) == ns else ' (every %d-th)' % (ns // want)}, best of {res['runs']} runs, "
                                         f"Peq build ({res['t_handle']:.3f} s) + kernel ({res['t_cal']:.3f} s)"}
    rec["_crcs"] = crcs
    del d_rows, d_packed, d_res, h_rows, h_res
    torch.cuda.empty_cache()
    return rec


def single_process_multi_device(ctx: Ctx, c5: dict, steps: int) -> dict | None:
    """Rank 0 alone drives devices 0..N-1 from ONE process, as the product does (bgsa_b200/host/aligner.c -g N; the
    reference: original/BGSA_AVX512/cal_mic.c:459-481): contiguous subject ranges, one bgsa_align_batch_submit per
    device, one pinned result buffer filled device-major.  The other ranks wait on the host (gloo) with idle GPUs."""
    torch, B, dist = ctx.torch, ctx.B, ctx.dist
    world = ctx.world
    out = None
    try:
        if ctx.rank == 0:
            out = _drive_all_devices(ctx, c5, steps)
    except Exception as exc:               # reported in the line; the other ranks must still be released
        out = {"error": f"{type(exc).__name__}: {exc}"}
    if dist is not None:
        dist.barrier(group=ctx.gloo)        # host-side wait: the other ranks' GPUs stay idle while rank 0 drives them
    return out


def _drive_all_devices(ctx: Ctx, c5: dict, steps: int) -> dict:
    torch, B = ctx.torch, ctx.B
    world = ctx.world
    if True:
        from concurrent.futures import ThreadPoolExecutor
        from bgsa_b200 import sharding
        wl = WORKLOADS["C5"]
        cfg = synth.CONFIGS["C5"]
        qlen, slen = cfg["qlen"], cfg["slen"]
        per = c5["config"]["subjects_per_gpu"]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=world) as ex:       # shard r = what rank r aligned in the C5 workload above
            parts = list(ex.map(lambda r: synth.make("C5", per, shard=r), range(world)))
        query = parts[0][0]
        total = per * world
        h_rows = torch.empty(total * (slen + 1), dtype=torch.uint8).pin_memory()
        rows = h_rows.numpy().reshape(total, slen + 1)
        for r, (_, s) in enumerate(parts):
            rows[r * per:(r + 1) * per] = s
        del parts
        h_res = torch.zeros(total, dtype=torch.int16).pin_memory()
        scores = h_res.numpy().reshape(1, total)
        log(f"single-process: {total} x {slen} bp subjects staged in {time.perf_counter() - t0:.1f} s")
        params = B.Params.default(wl["algo"], **wl["kw"])
        B.init_devices(world)
        ranges = [sharding.shard_range(total, g, world) for g in range(world)]

        def run(devices):
            for g in devices:
                first, cnt = ranges[g]
                B.align_batch_submit(params, query, rows, first, cnt, scores[:, first:first + cnt], device=g, slot=0)
            for g in devices:
                B.align_batch_wait(g, 0)

        def timed(devices, n):
            run(devices)                                        # warm-up (buffers, instance choice)
            t = time.perf_counter()
            for _ in range(n):
                run(devices)
            return (time.perf_counter() - t) / n

        t_one = timed([0], max(2, steps // 2))
        t_all = timed(list(range(world)), steps)
        cells_dev = float(qlen) * slen * per
        # device ranges are cut on tile boundaries (sharding.py), the ranks' shards are not: compare shard by shard
        crcs = [zlib.crc32(scores[:, r * per:(r + 1) * per].tobytes()) for r in range(world)]
        want = [c & 0xffffffff for c in c5["_crcs"]]
        out = {"workload": "C5 set, all ranks' shards concatenated", "devices": world, "subjects": total,
               "api": "one process: bgsa_align_batch_submit per device on contiguous ranges (bgsa_b200/sharding.py), "
                      "bgsa_align_batch_wait, one pinned device-major result buffer",
               "e2e": {"value": cells_dev * world / t_all / 1e9, "unit": "GCUPS", "ms_per_step": 1e3 * t_all,
                       "h2d_bytes_per_step": int(total * (slen + 1)), "d2h_bytes_per_step": int(total * 2)},
               "e2e_one_device": {"value": cells_dev / t_one / 1e9, "unit": "GCUPS", "ms_per_step": 1e3 * t_one,
                                  "note": "the same process driving device 0 alone on its range"},
               "speedup_over_one_device": (cells_dev * world / t_all) / (cells_dev / t_one),
               "device_ranges": [[int(f), int(c)] for f, c in ranges],
               "parity": {"crc32_per_shard": ["%08x" % c for c in crcs],
                          "shards_equal_to_rank_scores": int(sum(1 for a, b in zip(crcs, want) if a == b)),
                          "of": world,
                          "note": "rank r's scores of the C5 workload above were checked against the reference on a sample "
                                  "of its shard (workloads.C5.parity); the concatenated set holds rank r's shard at [r * per, (r + 1) * per)"}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--count", type=int, default=None, help="subjects per rank of the headline workload (default: the config's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--only-headline", action="store_true", help="skip the other configurations ('workloads') and the single-process run")
    ap.add_argument("--extra", default=None, help="comma-separated subset of the other configurations to run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    wl = WORKLOADS[args.workload]
    count = args.count or wl["count"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1 and args.impl == "ours":
        # convenience: re-launch ourselves one rank per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", "29517", __file__] + sys.argv[1:])
    emit = _claim_stdout()
    config = workload_config(args.workload, count)

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
            "statistic": "best of the runs (our arm reports the mean of its timed steps)",
            "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": res["cores"], "kind": res["kind"], "variant": res["variant"],
                             "sample": f"full workload ({subjects.shape[0]} subjects), best of {res['runs']} runs, Peq build + kernel",
                             "cal_only_gcups": res["gcups_cal"]},
            "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ---------------------------------------------------------------- our arm
    ctx = Ctx(args)
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    cache: dict = {}
    t_run = time.perf_counter()
    head = measure(ctx, args.workload, args.steps, args.warmup, count, cpu_seconds=10.0, data_cache=cache)
    extras = {}
    if not args.only_headline:
        names = [n for n in EXTRA_ORDER if n != args.workload]
        if args.extra is not None:
            names = [n for n in names if n in args.extra.split(",")]
        for n in names:
            steps = min(args.steps, WORKLOADS[n].get("max_steps", args.steps))
            try:
                extras[n] = measure(ctx, n, steps, 3, None, cpu_seconds=3.0, data_cache=cache)
            except Exception as exc:      # a failing side workload must not take the headline down; it is reported
                extras[n] = {"error": f"{type(exc).__name__}: {exc}"}
            if rank == 0:
                r = extras[n]
                log(f"{n}: " + (r["error"] if "error" in r else f"value {r['value']:.0f} GCUPS, e2e {r['e2e']['value']:.0f}, "
                                f"mismatches {r['parity']['mismatches']}/{r['parity']['subjects_compared']}") + f"  [{time.perf_counter() - t_run:.0f} s]")
    cache.clear()
    single = None
    if world > 1 and not args.only_headline:
        c5 = extras.get("C5") if args.workload != "C5" else head
        if c5 is not None and "error" not in c5:
            single = single_process_multi_device(ctx, c5, steps=min(args.steps, 5))
    clocks = sampler.stop() if rank == 0 else None
    int_peak2, _ = ctx.B.int_peak(ctx.dev)
    if rank != 0:
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()
        return

    for r in [head] + list(extras.values()):
        r.pop("_crcs", None)
    line = {
        "metric": "GCUPS", "value": head["value"], "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": config,
        "statistic": "mean of the timed steps (the reference arm reports the best of its runs)",
        "kernel": head["kernel"], "value_align_kernel_only": head["value_align_kernel_only"],
        "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
        "clocks": clocks, "sm_mhz_probe": ctx.sm_mhz_probe, "int_peak_probe_after": int_peak2 / 1e12,
        "host": {"numa_node_of_rank0": ctx.numa_node, "cores": ctx.host_cores},
        "git": _git_sha(),
    }
    if "two_kernel_path" in head:
        line["two_kernel_path"] = head["two_kernel_path"]
    if "cpu_baseline" in head:
        line["cpu_baseline"] = head["cpu_baseline"]
    line["parity"] = head["parity"]
    if extras:
        line["workloads"] = extras
    if single is not None:
        line["single_process"] = single
    line["bench_wall_s"] = time.perf_counter() - t_run
    emit(line)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
