"""Distils `ncu --page raw --csv` exports (tools/profile_all.sh) into profiles/: one compact CSV of the metrics
that evidence the design choices (ALU/FMA pipe utilisation, issue rate, DRAM bytes, occupancy, stall reasons)
and profiles/ncu_summary.json, which bench.py reads for `roofline.traffic`.
    python tools/ncu_summarize.py gpurun_out/r02prof r02 <git sha of the captured build>"""
import csv, json, sys
from pathlib import Path

src, tag = Path(sys.argv[1]), sys.argv[2]
git_sha = sys.argv[3] if len(sys.argv) > 3 else None      # commit the captured library was built from
ROOT = Path(__file__).resolve().parent.parent
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]
SUBJECTS = {"C2": 1_000_000, "C3": 10_000_000, "C3s": 10_000_000, "C4": 1_000_000, "C5": 32768, "myers150": 1_000_000, "C2np": 300_000,
            "myers5k": 16384, "C4_wavefront_K24_L2": 300_000}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}


def num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    return x * UNIT.get(unit, 1.0)


rows, summary = [], {}
for f in sorted(src.glob("full_*.raw.csv")):
    name = f.name[len("full_"):-len(".raw.csv")]
    wl, kern = name.split("_", 1)
    r = list(csv.reader(open(f)))
    hdr = next(i for i, x in enumerate(r) if "Kernel Name" in x)
    h, units, vals = r[hdr], r[hdr + 1], r[hdr + 2]
    d = {n: (vals[i], units[i]) for i, n in enumerate(h)}
    kname = d.get("Kernel Name", ("?", ""))[0]
    for m in KEEP:
        if m in d:
            rows.append([tag, wl, kern, m, d[m][0], d[m][1]])
    rd, wr = num(*d["dram__bytes_read.sum"]), num(*d["dram__bytes_write.sum"])
    dur = num(*d["gpu__time_duration.sum"])
    nsub = SUBJECTS.get(f"{wl}_{kern}", SUBJECTS[wl])
    entry = {"kernel": kname[:140], "subjects_in_capture": nsub, "dram_bytes_per_launch": rd + wr,
             "dram_bytes_per_subject": (rd + wr) / nsub, "duration_s_under_ncu": dur,
             "alu_pipe_pct": float(d["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"][0]),
             "fma_pipe_pct": float(d["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"][0]),
             "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
             "dram_pct_of_peak": float(d["dram__cycles_active.avg.pct_of_peak_sustained_elapsed"][0]),
             "registers": int(float(d["launch__registers_per_thread"][0])),
             "source": f"profiles/{tag}_ncu_metrics.csv (ncu --set full --clock-control none, tools/profile_all.sh)",
             "captured_at_git": git_sha}
    stalls = {m.split("issue_stalled_")[1].split("_per_issue")[0]: float(d[m][0]) for m in KEEP if "issue_stalled" in m and m in d}
    entry["stalls_per_issue"] = stalls
    main = kern in ("align_rows_kernel", "banded_kernel", "align_kernel")
    summary[wl if main else f"{wl}_{kern}"] = entry
out = ROOT / "profiles" / f"{tag}_ncu_metrics.csv"
with open(out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["round", "workload", "kernel", "metric", "value", "unit"])
    w.writerows(rows)
(ROOT / "profiles" / "ncu_summary.json").write_text(json.dumps(summary, indent=1))
for k, v in summary.items():
    print(f"{k:14s} ALU {v['alu_pipe_pct']:5.1f}%  FMA {v['fma_pipe_pct']:5.1f}%  issue {v['issue_active_pct']:5.1f}%  DRAM {v['dram_pct_of_peak']:5.1f}%  "
          f"{v['dram_bytes_per_subject']:8.1f} B/subject  regs {v['registers']}  {v['duration_s_under_ncu']*1e3:8.3f} ms")
