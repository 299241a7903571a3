#!/bin/bash
# One-GPU profiling pass (run under gpurun): launch lists + one full ncu capture per hot kernel,
# at the bench configurations.  Usage: tools/gpu_profile.sh <round-tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
B="--no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_${TAG}_shkadov.csv \
    python bench.py --steps 20 --warmup 3 $B > $OUT/launches_${TAG}_shkadov.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shkadov_kernel -s 4 -c 1 -f -o $OUT/prof_${TAG}_shkadov \
    python bench.py --steps 6 --warmup 3 $B > $OUT/prof_${TAG}_shkadov.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file $OUT/launches_${TAG}_rayleigh.csv \
    python bench.py --env rayleigh --steps 3 --warmup 3 $B > $OUT/launches_${TAG}_rayleigh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mac_ -s 4 -c 1 -f -o $OUT/prof_${TAG}_rayleigh \
    python bench.py --env rayleigh --steps 3 --warmup 3 $B > $OUT/prof_${TAG}_rayleigh.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 10 --csv --log-file $OUT/launches_${TAG}_mixing.csv \
    python bench.py --env mixing --steps 2 --warmup 3 $B > $OUT/launches_${TAG}_mixing.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mac_ -s 4 -c 1 -f -o $OUT/prof_${TAG}_mixing \
    python bench.py --env mixing --steps 2 --warmup 3 $B > $OUT/prof_${TAG}_mixing.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:shkadov_kernel -s 4 -c 1 -f -o $OUT/prof_${TAG}_shkadov_separable \
    python bench.py --env shkadov_separable --steps 6 --warmup 3 $B > $OUT/prof_${TAG}_shkadov_separable.log 2>&1
ls -la $OUT
