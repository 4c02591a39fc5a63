for rep in 1 2; do
  for lib in ab/new1.so ab/new3.so; do
    echo "== $lib"
    SGFHE_CUDA_LIB=$PWD/$lib SGFHE_PHASE_TIMING=1 timeout 300 python bench.py --n 512 --batch 1184 --steps 2 --warmup 1 --no-cpu 2>&1 \
      | grep -E "sgfhe phase|\"value\"" | tail -9 | sed -e 's/.*"value": \([0-9.]*\).*"verified": \([a-z]*\).*/gates_per_s \1 verified \2/' | cut -c1-100
  done
done
