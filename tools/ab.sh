#!/bin/bash
# A/B two builds of libsgfhe_cuda.so on the same box: tools/ab.sh <base.so> <new.so> [batch]
# prints gates/s and the phase split for each, alternating twice to expose drift.
B=${3:-148}
for rep in 1 2; do
  for lib in "$1" "$2"; do
    echo "== $lib"
    SGFHE_CUDA_LIB=$PWD/$lib SGFHE_PHASE_TIMING=1 timeout 300 python bench.py --n 1024 --batch $B --steps 1 --warmup 1 --no-cpu 2>&1 \
      | grep -E "sgfhe phase|\"value\"" | sed -e 's/.*"value": \([0-9.]*\).*"verified": \([a-z]*\).*/gates_per_s \1 verified \2/' | cut -c1-100
  done
done
