# Round-2 closing runs: kernel sweep (same operand bytes at every ring degree, Params(2048) included), Params(2048) gates,
# and a 2-rank torchrun sanity run of the default bench (needs gpurun --gpus 2; skipped on one GPU).
set -x
python bench_kernels.py > gpurun_out/bench_kernels_r02.jsonl 2> gpurun_out/bench_kernels_r02.err; cut -c1-260 gpurun_out/bench_kernels_r02.jsonl | head -8
python bench.py --n 2048 --batch 296 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_r02_p2048.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r02_p2048.json
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
