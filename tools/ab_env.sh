#!/bin/bash
# A/B two environment settings of the same libsgfhe_cuda.so on one box: tools/ab_env.sh "<env A>" "<env B>" [batch]
# prints gates/s and the phase split for each, alternating twice to expose drift.
B=${3:-148}
for rep in 1 2; do
  for e in "$1" "$2"; do
    echo "== env: $e"
    env $e SGFHE_PHASE_TIMING=1 timeout 300 python bench.py --n 1024 --batch $B --steps 1 --warmup 1 --no-cpu 2>&1 \
      | grep -E "sgfhe phase|\"value\"" | sed -e 's/.*"value": \([0-9.]*\).*"verified": \([a-z]*\).*/gates_per_s \1 verified \2/' | cut -c1-100
  done
done
