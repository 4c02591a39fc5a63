# Final checks of round 2: GPU parity suite, smoke, default bench, randomised-mode bench, Params(2048) and Params(512) lines.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; cut -c1-300 gpurun_out/bench_r02_final.json
python bench.py --steps 3 --warmup 3 --no-cpu --rng-seed 7 > gpurun_out/bench_r02_rng.json 2> gpurun_out/bench_r02_rng.err; cut -c1-300 gpurun_out/bench_r02_rng.json
python bench.py --n 512 --batch 4096 --steps 2 --warmup 2 --no-cpu > gpurun_out/bench_r02_p512_b4096.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r02_p512_b4096.json
python bench.py --n 2048 --batch 296 --steps 1 --warmup 1 --no-cpu > gpurun_out/bench_r02_p2048.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r02_p2048.json
