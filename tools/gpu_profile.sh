#!/bin/bash
# One-GPU profiling pass (run under gpurun): launch list of the default bench command + one full ncu capture per hot
# kernel at the bench configurations.  Usage: tools/gpu_profile.sh <round-tag>
TAG=${1:-r2}
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
  # -s: skip the reset launch(es) and the warm-up steps
  S=4; [ $e = shkadov ] && S=6; [ $e = shkadov_separable ] && S=6
  ncu --set full --clock-control none --import-source on -k $R -s $S -c 1 -f -o $OUT/prof_${TAG}_$e \
      python bench.py --env $e --steps $K --warmup 3 $B > $OUT/prof_${TAG}_$e.log 2>&1
done
ls -la $OUT | tail -12
