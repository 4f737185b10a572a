#!/bin/bash
# Profiling pass for one round (run under gpurun, 1 GPU):  bash tools/profile_all.sh <out dir under gpurun_out>
#   1. launch list of the default bench command (every launch with its device time; cold-cache, serialised)
#   2. one `ncu --set full` capture of the dominant kernel of every workload, of a forced wavefront instance and of the
#      two-kernel path (pack + align on packed tiles)
# Each ncu run follows a plain run of the same command that exited 0 (B200_PROFILING.md).
set -u
OUT=gpurun_out/${1:-prof}
mkdir -p "$OUT"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > "$OUT/bench_plain.log" 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench_C2.csv" $BENCH > "$OUT/bench_under_ncu.log" 2>&1
# workload, subjects, kernel regex, tag suffix, extra environment (BGSA_FORCE_KL pins an instance)
while read -r wl n kern tag envs; do
    [ -z "$wl" ] && continue
    env $envs python tools/prof_one.py $wl $n 2 > "$OUT/plain_${wl}_$tag.log" 2>&1 &&
    env $envs ncu --set full --clock-control none --import-source on -k regex:$kern -s 1 -c 1 -f -o "$OUT/full_${wl}_$tag" python tools/prof_one.py $wl $n 2 > "$OUT/ncu_${wl}_$tag.log" 2>&1
    tail -1 "$OUT/ncu_${wl}_$tag.log"
    # gpurun brings back at most 64 MiB: keep the raw metric page of every capture, the report itself only for C2
    ncu -i "$OUT/full_${wl}_$tag.ncu-rep" --page raw --csv > "$OUT/full_${wl}_$tag.raw.csv" 2>/dev/null
    ncu -i "$OUT/full_${wl}_$tag.ncu-rep" --page details --csv > "$OUT/full_${wl}_$tag.details.csv" 2>/dev/null
    if [ "${wl}_$tag" != "C2_align_rows_kernel" ]; then rm -f "$OUT/full_${wl}_$tag.ncu-rep"; fi
done <<'SPECS'
C2 1000000 align_rows_kernel align_rows_kernel
C3 10000000 banded_kernel banded_kernel
C3s 10000000 banded_kernel banded_kernel
C4 1000000 align_kernel align_kernel
C5 32768 align_kernel align_kernel
myers150 1000000 align_rows_kernel align_rows_kernel
C2np 300000 align_rows_kernel align_rows_kernel
myers5k 16384 align_kernel align_kernel
C4 300000 align_kernel wavefront_K24_L2 BGSA_FORCE_KL=24,2
C2 1000000 align_kernel packed_path BGSA_NO_ROWS_KERNEL=1
C2 1000000 pack_stream pack_stream BGSA_NO_ROWS_KERNEL=1
SPECS
ls -la "$OUT"
