#!/bin/bash
# All bench lines of a round (run under gpurun): tools/bench_all.sh <round-tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
for e in shkadov rayleigh mixing shkadov_separable sloshing burgers lorenz; do
  python bench.py --env $e > $OUT/${TAG}_bench_$e.json 2> $OUT/${TAG}_bench_$e.err
done
python bench.py --impl reference > $OUT/${TAG}_bench_reference_shkadov.json 2>/dev/null
python bench.py --impl reference --env rayleigh > $OUT/${TAG}_bench_reference_rayleigh.json 2>/dev/null
for f in $OUT/${TAG}_bench_*.json; do echo "$f: $(cut -c1-260 $f)"; done
