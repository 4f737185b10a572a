#!/bin/bash
# Profiling pass for one round (run under gpurun, 1 GPU):  bash tools/profile_all.sh <out dir under gpurun_out>
#   1. launch list of the default bench command (every launch with its device time; cold-cache, serialised)
#   2. one `ncu --set full` capture of the dominant kernel of every workload + the pack kernel
# Each ncu run follows a plain run of the same command that exited 0 (B200_PROFILING.md).
set -u
OUT=gpurun_out/${1:-prof}
mkdir -p "$OUT"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > "$OUT/bench_plain.log" 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench_C2.csv" $BENCH > "$OUT/bench_under_ncu.log" 2>&1
for spec in "C2 1000000 align_kernel" "C3 10000000 banded_kernel" "C4 1000000 align_kernel" "C5 32768 align_kernel" "myers150 1000000 align_kernel" "C4 1000000 pack_stream" "C2 1000000 pack_stream"; do
    set -- $spec
    python tools/prof_one.py $1 $2 2 > "$OUT/plain_$1_$3.log" 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -f -o "$OUT/full_$1_$3" python tools/prof_one.py $1 $2 2 > "$OUT/ncu_$1_$3.log" 2>&1
    tail -1 "$OUT/ncu_$1_$3.log"
    # gpurun brings back at most 64 MiB: keep the raw metric page of every capture, the report itself only for C2
    ncu -i "$OUT/full_$1_$3.ncu-rep" --page raw --csv > "$OUT/full_$1_$3.raw.csv" 2>/dev/null
    ncu -i "$OUT/full_$1_$3.ncu-rep" --page details --csv > "$OUT/full_$1_$3.details.csv" 2>/dev/null
    if [ "$1_$3" != "C2_align_kernel" ]; then rm -f "$OUT/full_$1_$3.ncu-rep"; fi
done
ls -la "$OUT"
