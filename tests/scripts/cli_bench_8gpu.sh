#!/bin/bash
# TEST INFRASTRUCTURE (runs the reference aligner of oracle/_ref as the checker).
# BASELINE config 5 at the CLI: BitPAl 2/-3/-5, 5 kbp query x N x 5 kbp subjects, one process driving all GPUs of the
# box (aligner -g <n>), result file compared byte for byte with the reference aligner's.
#   bash tests/scripts/cli_bench_8gpu.sh <out dir under gpurun_out> [subjects]
set -u
OUT=gpurun_out/${1:-cli8}; N=${2:-200000}
mkdir -p "$OUT" /tmp/cli8
python - <<PY
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import synth, refutil as R
q, s = synth.make("C5", $N)
R.write_rows("/tmp/cli8/q.txt", q); R.write_rows("/tmp/cli8/s.txt", s)
PY
NG=$(nvidia-smi -L | wc -l)
run() { local t0=$(date +%s%N); "$@" > /tmp/cli8/last.log 2>&1; local rc=$?; local t1=$(date +%s%N); echo "rc=$rc wall=$(( (t1 - t0) / 1000000 )) ms :: $*"; grep -E "GCUPS|time " /tmp/cli8/last.log | tr '\n' ';'; echo; }
{
echo "== C5 at the CLI: 5 kbp x $N x 5 kbp, $NG GPUs present"
run bgsa_b200/aligner -v -g $NG -a bitpal -q /tmp/cli8/q.txt -d /tmp/cli8/s.txt -f /tmp/cli8/ours_g.bin
run bgsa_b200/aligner -v -g $NG -a bitpal -q /tmp/cli8/q.txt -d /tmp/cli8/s.txt -f /tmp/cli8/ours_g.bin
run bgsa_b200/aligner -v -g 1 -a bitpal -q /tmp/cli8/q.txt -d /tmp/cli8/s.txt -f /tmp/cli8/ours_1.bin
[ -x oracle/_ref/aligner_bitpal_avx512 ] && run oracle/_ref/aligner_bitpal_avx512 -q /tmp/cli8/q.txt -d /tmp/cli8/s.txt -f /tmp/cli8/ref.bin
cmp /tmp/cli8/ours_g.bin /tmp/cli8/ref.bin && echo "-g $NG result file identical to the reference aligner's"
cmp /tmp/cli8/ours_1.bin /tmp/cli8/ref.bin && echo "-g 1 result file identical to the reference aligner's"
bgsa_b200/convert -r /tmp/cli8/ours_g.bin -o /tmp/cli8/ours_g.txt > /dev/null && oracle/_ref/convert_int16 -r /tmp/cli8/ref.bin -o /tmp/cli8/ref.txt > /dev/null && cmp /tmp/cli8/ours_g.txt /tmp/cli8/ref.txt && echo "converted text identical ($NG-device .info against the reference's 1-device .info)"
} 2>&1 | tee "$OUT/cli_bench_8gpu.log"
