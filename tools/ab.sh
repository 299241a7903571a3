#!/bin/bash
# A/B builds of libbeacon_b200.so on the same GPU box: tools/ab.sh libA.so libB.so ... -- bench args
LIBS=(); while [ "$1" != "--" ] && [ -n "$1" ]; do LIBS+=($1); shift; done; shift
L=beacon_b200/lib/libbeacon_b200.so
cp $L /tmp/lib_keep.so
for rep in 1 2; do
  for v in "${LIBS[@]}"; do cp $v $L; echo -n "$v: "; python bench.py "$@" --no-cpu-baseline --no-e2e 2>&1 | grep -o '"value": [0-9.]*' | head -1; done
done
cp /tmp/lib_keep.so $L
