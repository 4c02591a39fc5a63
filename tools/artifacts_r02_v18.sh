# Final artifacts of round 2 for kernel v18 (CRT pre-scaling folded into the key words): GPU parity suite, default bench,
# reference arm, launch list, one ncu --set full capture of a 148-gate launch.  Outputs in gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_v18.json 2> gpurun_out/bench_r02_v18.err; cut -c1-400 gpurun_out/bench_r02_v18.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_v18_ref.json 2> gpurun_out/bench_r02_v18_ref.err; cut -c1-300 gpurun_out/bench_r02_v18_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_v18.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ncu18_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bootstrap_kernel_v4 -c 1 -f -o gpurun_out/r2_v18 python bench.py --n 1024 --batch 148 --steps 1 --warmup 0 --no-cpu > gpurun_out/r2_ncu18.log 2>&1
ls -la gpurun_out/r2_v18.ncu-rep
