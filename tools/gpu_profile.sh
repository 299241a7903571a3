#!/bin/bash
# One-GPU profiling pass (run under gpurun): launch list of the default bench command + one full ncu capture per hot
# kernel at the bench configurations, summarised ON THE BOX (gpurun brings back at most 64 MiB; a report is ~13 MB).
# Usage: tools/gpu_profile.sh <round-tag> [keep-report-of-env]
TAG=${1:-r2}
KEEP=${2:-none}
OUT=gpurun_out
mkdir -p $OUT
B="--no-cpu-baseline --no-e2e --extras none"
# launch list of the SAME command the driver times (all workloads of the default line; per-launch times are cold-cache
# and serialised: only each kernel's share of its step is comparable)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}_default.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/launches_${TAG}_default.log 2>&1
for e in shkadov rayleigh mixing shkadov_separable; do
  K=3; [ $e = shkadov ] && K=6; [ $e = shkadov_separable ] && K=6; [ $e = mixing ] && K=2
  R="regex:mac_"; [ $e = shkadov ] && R="regex:shkadov_kernel"; [ $e = shkadov_separable ] && R="regex:shkadov_kernel"
  S=4; [ $e = shkadov ] && S=6; [ $e = shkadov_separable ] && S=6      # skip the reset launch(es) and the warm-up steps
  ncu --set full --clock-control none --import-source on -k $R -s $S -c 1 -f -o /tmp/prof_${TAG}_$e \
      python bench.py --env $e --steps $K --warmup 3 $B > $OUT/prof_${TAG}_$e.log 2>&1
  python tools/summarize_ncu.py /tmp/prof_${TAG}_$e.ncu-rep $OUT/${TAG}_ncu_$e.json
  python tools/ncu_lines.py /tmp/prof_${TAG}_$e.ncu-rep 40 > $OUT/${TAG}_ncu_lines_$e.txt 2>&1
  [ "$KEEP" = "$e" ] && cp /tmp/prof_${TAG}_$e.ncu-rep $OUT/
done
ls -la $OUT | tail -16
