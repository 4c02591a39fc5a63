#!/bin/bash
# A/B two builds on the smaller parameter sets: tools/ab_small.sh <base.so> <new.so>
for n in 64 128 256; do
  for rep in 1 2; do
    for lib in "$1" "$2"; do
      echo -n "n=$n $lib: "
      SGFHE_CUDA_LIB=$PWD/$lib timeout 300 python bench.py --n $n --batch 9472 --steps 3 --warmup 2 --no-cpu 2>/dev/null \
        | sed -e 's/.*"value": \([0-9.]*\).*"verified": \([a-z]*\).*/gates_per_s \1 verified \2/' | cut -c1-100
    done
  done
done
