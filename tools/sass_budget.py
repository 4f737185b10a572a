#!/usr/bin/env python
"""sass_budget.py -- per-pipe instruction budget of the hot loops of the alignment kernels, read from the SASS of
the built library (no GPU needed).

    python tools/sass_budget.py [--lib bgsa_b200/libbgsa_b200.so] [--out profiles/r02_sass] [--json FILE] [name ...]

For every named kernel (table KERNELS below; default: all) it
  * finds the kernel in `cuobjdump -sass`, demangled with cu++filt,
  * finds its innermost loops (backward branches that contain no other backward branch),
  * picks the hot loop = the innermost loop with the most integer-pipe work that loads the query's match masks
    (LDS) -- the unrolled column loop of align_kernel, the 32-row block of banded_kernel,
  * counts its instructions per issue pipe (ALU: LOP3/IADD3/SHF/LEA/PRMT/SEL/ISETP/MOV/...; FMA: IMAD*/FFMA...;
    LSU: LDS/LDG/STS/...; control, uniform datapath),
  * divides by the DP columns (or band rows) one trip computes and by the cells of a column, giving
    `ops_per_cell_sass` (ALU pipe) -- the denominator bench.py uses for roofline.frac_sass,
  * checks the carry chains: every add chain of K words must be K IADD3/IADD3.X with no other carry writer between
    its links (SURVEY 8d asks for I confirmed from SASS; VERDICT r01 "fragility" asks for this guard),
  * writes the loop listing to <out>/<name>.sass and a summary table to <out>/summary.md (+ --json).

Pipes follow B300_MICROARCH.md "Pipe rates": IMAD/FFMA on the fma pipe, IADD3/LOP3/SHF/PRMT/LEA on the alu pipe, each
one warp-instruction per 2 clocks per SM sub-partition.
"""
from __future__ import annotations

import argparse
import json
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

# name -> (regex on the demangled kernel name, columns/rows per trip or None = last template argument (UNROLL),
#          query words per column (K), cells per column = query length of the workload)
KERNELS = {
    "C2_bitpal_packed_K5":      (r"align_kernel<bgsa::BitpalPacked<bgsa::Scheme<2, -3, -5>, 5, 0>, 1,", None, 5, 150),
    "C2_rows_bitpal_packed_K5": (r"align_rows_kernel<bgsa::BitpalPacked<bgsa::Scheme<2, -3, -5>, 5, 0>, 128,", None, 5, 150),
    "myers150_rows_K5":         (r"align_rows_kernel<bgsa::MyersAlgo<5, 0>, 128,", None, 5, 150),
    "C2np_rows_K5":             (r"align_rows_kernel<bgsa::BitpalNonPacked<bgsa::Scheme<2, -3, -5>, 5>, 128,", None, 5, 150),
    "C5_bitpal_packed_K10_L16": (r"align_kernel<bgsa::BitpalPacked<bgsa::Scheme<2, -3, -5>, 10, 0>, 16,", None, 10, 312.5),
    "C4_myers_semi_K32":        (r"align_kernel<bgsa::MyersAlgo<32, 1>, 1,", None, 32, 1000),
    "myers150_K5":              (r"align_kernel<bgsa::MyersAlgo<5, 0>, 1,", None, 5, 150),
    "myers5k_K20_L8":           (r"align_kernel<bgsa::MyersAlgo<20, 0>, 8,", None, 20, 625),
    "bitpal_nonpacked_150":     (r"align_kernel<bgsa::BitpalNonPacked<bgsa::Scheme<2, -3, -5>, 5>, 1,", None, 5, 150),
    # banded: the hot region is the fully unrolled 32-row block (straight-line code inside the block loop); "cells" of a
    # row = the subject length (nominal q x s cells, the reference's GCUPS convention, banded/BGSA_CPU/cal_cpu.c:469)
    "C3_banded_fused":          (r"banded_kernel<0, 0, 1,", -32, 1, 100),
    "C3_banded_packed":         (r"banded_kernel<0, 0, 0,", -32, 1, 100),
}

ALU = ("LOP3", "IADD3", "IADD", "SHF", "LEA", "PRMT", "SEL", "ISETP", "MOV", "IMNMX", "VIMNMX", "FMNMX", "BMSK", "SGXT", "LOP",
       "PLOP3", "IABS", "VABSDIFF", "FSETP", "FSEL", "I2I", "CS2R", "P2R", "R2P")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HMUL2", "HADD2", "IDP", "IMUL")
XU = ("POPC", "FLO", "BREV", "MUFU", "I2F", "F2I", "I2FP", "F2FP")
LSU = ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "ATOM", "ATOMS", "ATOMG", "RED", "LDSM", "LDC", "UBLKCP", "SYNCS")
CTRL = ("BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "WARPSYNC", "BAR", "NOP", "YIELD", "BRX", "JMP", "BREAK", "NANOSLEEP", "DEPBAR",
        "ERRBAR", "MEMBAR", "FENCE", "CCTL")
WARP = ("SHFL", "VOTE", "MATCH", "REDUX", "S2R", "S2UR", "R2UR", "ELECT")


def pipe_of(op: str) -> str:
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UBLKCP",) and len(base) > 1 and base[1:] in (
            "MOV", "IADD3", "LOP3", "SHF", "LEA", "ISETP", "IMAD", "SEL", "FLO", "POPC", "PRMT", "BMSK", "SGXT", "PLOP3", "LDC", "P2UR",
            "IADD", "BREV", "R2UR", "F2FP", "CLEA", "MEMBAR"):
        return "uniform"
    for names, pipe in ((FMA, "fma"), (ALU, "alu"), (XU, "xu"), (LSU, "lsu"), (CTRL, "ctrl"), (WARP, "warp")):
        if base in names:
            return pipe
    return "other"


INSTR_RE = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);")


def sass_functions(lib: Path):
    txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    funcs, cur, name = {}, None, None
    for line in txt.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            name = m.group(1)
            cur = funcs.setdefault(name, [])
            continue
        if cur is None:
            continue
        m = INSTR_RE.match(line)
        if m:
            cur.append((int(m.group(1), 16), m.group(2), m.group(3), line.rstrip()))
    names = list(funcs)
    dem = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
    # cu++filt writes template arguments as (int)32 / (bool)1: drop the casts
    return {re.sub(r"\((?:int|bool)\)", "", d): funcs[n] for n, d in zip(names, dem)}


def innermost_loops(instrs):
    addr_index = {a: i for i, (a, _, _, _) in enumerate(instrs)}
    back = []
    for i, (a, op, args, _) in enumerate(instrs):
        if op.split(".")[0] == "BRA":
            m = re.search(r"0x([0-9a-f]+)", args)
            if m:
                t = int(m.group(1), 16)
                if t <= a and t in addr_index:
                    back.append((addr_index[t], i))
    inner = [(s, e) for (s, e) in back if not any((s2 >= s and e2 <= e and (s2, e2) != (s, e)) for (s2, e2) in back)]
    return inner


def carry_chain_report(body):
    """Add chains in SASS: a chain starts with an IADD3 that writes a carry predicate (`IADD3 R, P0, PT, a, b, RZ`) and
    continues with IADD3.X links (carry in and out); ptxas turns a last link whose carry-out is dead into IMAD.X (FMA
    pipe).  Returns (starts, links) of the loop body."""
    starts = links = 0
    for _, op, args, _ in body:
        base = op.split(".")[0]
        if base == "IADD3" and ".X" not in op:
            parts = [a.strip() for a in args.split(",")]
            if len(parts) > 1 and re.fullmatch(r"P\d", parts[1]):
                starts += 1
        elif (base == "IADD3" and ".X" in op) or (base == "IMAD" and ".X" in op):
            links += 1
    return starts, links


def analyse(name, spec, funcs, outdir: Path | None):
    rx, per_trip, K, cells = spec
    cands = [(d, f) for d, f in funcs.items() if re.search(rx, d)]
    if not cands:
        return {"kernel": None, "error": f"no kernel matches {rx}"}
    dem, instrs = cands[0]
    if per_trip is None:
        m = re.search(r",\s*(\d+)>\(", dem)
        per_trip = int(m.group(1)) if m else 1
    best = None
    regions = innermost_loops(instrs)
    if per_trip < 0:            # straight-line mode: maximal runs without a branch or a branch target inside
        per_trip = -per_trip
        targets = set()
        for _, op, a, _ in instrs:
            if op.split(".")[0] in ("BRA", "BSSY"):
                m = re.search(r"0x([0-9a-f]+)", a)
                if m:
                    targets.add(int(m.group(1), 16))
        regions, start = [], 0
        for i, (addr, op, _, _) in enumerate(instrs):
            if addr in targets and i > start:
                regions.append((start, i - 1)); start = i
            if op.split(".")[0] in ("BRA", "EXIT", "BSYNC", "BSSY", "CALL", "RET"):
                if i - 1 >= start:
                    regions.append((start, i - 1))
                start = i + 1
    for s, e in regions:
        body = instrs[s:e + 1]
        pipes = Counter(pipe_of(op) for _, op, _, _ in body)
        has_lds = any(op.startswith("LDS") or op.startswith("LDG") for _, op, _, _ in body)
        score = (pipes["alu"] + pipes["fma"]) * (2 if has_lds else 1)
        if best is None or score > best[0]:
            best = (score, s, e, body, pipes)
    if best is None:
        return {"kernel": dem, "error": "no loop found"}
    _, s, e, body, pipes = best
    ops = Counter(op for _, op, _, _ in body)
    res = {
        "kernel": dem.split("(")[0],
        "loop": f"0x{instrs[s][0]:04x}..0x{instrs[e][0]:04x}",
        "instructions_per_trip": len(body),
        "columns_per_trip": per_trip,
        "words_per_column": K,
        "pipes_per_trip": dict(pipes),
        "alu_per_column": pipes["alu"] / per_trip,
        "fma_per_column": pipes["fma"] / per_trip,
        "issue_per_column": len(body) / per_trip,
        "alu_per_word_column": pipes["alu"] / per_trip / K,
        "ops_per_cell_sass": pipes["alu"] / per_trip / cells,
        "top_ops": dict(ops.most_common(12)),
        "carry_chain_starts_per_column": carry_chain_report(body)[0] / per_trip,
        "carry_chain_links_per_column": carry_chain_report(body)[1] / per_trip,
    }
    if outdir is not None:
        outdir.mkdir(parents=True, exist_ok=True)
        with open(outdir / f"{name}.sass", "w") as f:
            f.write(f"// {dem}\n// hot loop {res['loop']}: {len(body)} instructions, {per_trip} column(s) per trip\n")
            f.write("// per pipe: " + ", ".join(f"{k} {v}" for k, v in sorted(pipes.items())) + "\n")
            for _, _, _, line in body:
                f.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", line) + "\n")
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=str(ROOT / "bgsa_b200" / "libbgsa_b200.so"))
    ap.add_argument("--out", default=None, help="directory for the loop listings + summary.md")
    ap.add_argument("--json", default=None)
    ap.add_argument("names", nargs="*")
    args = ap.parse_args()
    funcs = sass_functions(Path(args.lib))
    outdir = Path(args.out) if args.out else None
    names = args.names or list(KERNELS)
    results = {}
    for n in names:
        results[n] = analyse(n, KERNELS[n], funcs, outdir)
    hdr = "| kernel | trip instr | col/trip | ALU/col | FMA/col | LSU/col | ALU per word-col | ops/cell (ALU, SASS) | add chains per column: starts + links |"
    lines = [hdr, "|---|---|---|---|---|---|---|---|---|"]
    for n, r in results.items():
        if "error" in r:
            lines.append(f"| {n} | {r['error']} |")
            continue
        p = r["pipes_per_trip"]
        lines.append(f"| {n} | {r['instructions_per_trip']} | {r['columns_per_trip']} | {r['alu_per_column']:.1f} | {r['fma_per_column']:.1f} | "
                     f"{p.get('lsu', 0) / r['columns_per_trip']:.1f} | {r['alu_per_word_column']:.2f} | {r['ops_per_cell_sass']:.3f} | "
                     f"{r['carry_chain_starts_per_column']:.1f} + {r['carry_chain_links_per_column']:.1f} |")
    table = "\n".join(lines)
    print(table)
    if outdir is not None:
        outdir.mkdir(parents=True, exist_ok=True)
        (outdir / "summary.md").write_text("# SASS budget of the hot loops (tools/sass_budget.py)\n\n" + table + "\n")
    if args.json:
        Path(args.json).write_text(json.dumps(results, indent=1) + "\n")


if __name__ == "__main__":
    main()
