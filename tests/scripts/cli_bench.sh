#!/bin/bash
# TEST INFRASTRUCTURE (runs the reference binaries of oracle/_ref as the checker).
# End-to-end CLI comparison on files (run under gpurun: bash tests/scripts/cli_bench.sh <out>): our aligner vs the reference's aligner on the same
# synthetic C2 / C3 files, byte-compare of the result files, wall times.
set -u
OUT=gpurun_out/${1:-cli}
mkdir -p "$OUT" /tmp/cli
python - <<'PY'
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, "tests")
import synth, refutil as R
for name, n in (("C2", 1_000_000), ("C3", 4_000_000), ("C5", 40_000)):
    q, s = synth.make(name, n)
    R.write_rows(f"/tmp/cli/{name}_q.txt", q); R.write_rows(f"/tmp/cli/{name}_s.txt", s)
PY
run() { local t0=$(date +%s%N); "$@" > /tmp/cli/last.log 2>&1; local rc=$?; local t1=$(date +%s%N); echo "rc=$rc wall=$(( (t1 - t0) / 1000000 )) ms :: $*"; grep -E "GCUPS|time " /tmp/cli/last.log | tr '\n' ';'; echo; }
{
echo "== C2: BitPAl 2/-3/-5, 1 x 1M x 150bp"
run bgsa_b200/aligner -v -a bitpal -q /tmp/cli/C2_q.txt -d /tmp/cli/C2_s.txt -f /tmp/cli/ours_c2.bin
run bgsa_b200/aligner -v -a bitpal -q /tmp/cli/C2_q.txt -d /tmp/cli/C2_s.txt -f /tmp/cli/ours_c2.bin
[ -x oracle/_ref/aligner_bitpal_avx512 ] && run oracle/_ref/aligner_bitpal_avx512 -q /tmp/cli/C2_q.txt -d /tmp/cli/C2_s.txt -f /tmp/cli/ref_c2.bin
cmp /tmp/cli/ours_c2.bin /tmp/cli/ref_c2.bin && echo "C2 result files identical"
echo "== C3: banded e=5, 1 x 4M x 100bp"
run bgsa_b200/aligner -v -a banded -k 5 -q /tmp/cli/C3_q.txt -d /tmp/cli/C3_s.txt -f /tmp/cli/ours_c3.bin
[ -x oracle/_ref/aligner_banded_cpu ] && run oracle/_ref/aligner_banded_cpu -k 5 -q /tmp/cli/C3_q.txt -d /tmp/cli/C3_s.txt -f /tmp/cli/ref_c3.bin
cmp /tmp/cli/ours_c3.bin /tmp/cli/ref_c3.bin && echo "C3 result files identical"
echo "== C5 slice: BitPAl 2/-3/-5, 5 kbp query x 40k x 5 kbp"
NG=$(nvidia-smi -L | wc -l)
run bgsa_b200/aligner -v -a bitpal -q /tmp/cli/C5_q.txt -d /tmp/cli/C5_s.txt -f /tmp/cli/ours_c5.bin
[ "$NG" -gt 1 ] && run bgsa_b200/aligner -v -g $NG -a bitpal -q /tmp/cli/C5_q.txt -d /tmp/cli/C5_s.txt -f /tmp/cli/ours_c5_g.bin
[ -x oracle/_ref/aligner_bitpal_avx512 ] && run oracle/_ref/aligner_bitpal_avx512 -q /tmp/cli/C5_q.txt -d /tmp/cli/C5_s.txt -f /tmp/cli/ref_c5.bin
cmp /tmp/cli/ours_c5.bin /tmp/cli/ref_c5.bin && echo "C5 result files identical"
[ "$NG" -gt 1 ] && cmp /tmp/cli/ours_c5_g.bin /tmp/cli/ref_c5.bin && echo "C5 multi-GPU result file identical (one query: payload order is device independent)"
} 2>&1 | tee "$OUT/cli_bench.log"
