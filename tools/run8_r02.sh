# 8-GPU runs of round 2 (one box): weak and strong scaling at Params(1024), the depth.jl chain at Params(512) with
# all-gathered and with rank-local wiring.  Artifacts land in gpurun_out/, copies of the JSON lines in profiles/.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
if [ "$1" != "depth-only" ]; then
$TR bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r02_8gpu_weak.json 2> gpurun_out/bench_r02_8gpu_weak.err; cut -c1-300 gpurun_out/bench_r02_8gpu_weak.json
$TR bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --scaling strong --batch 4096 > gpurun_out/bench_r02_8gpu_strong.json 2> gpurun_out/bench_r02_8gpu_strong.err; cut -c1-300 gpurun_out/bench_r02_8gpu_strong.json
fi
$TR bench.py --gpus 8 --workload depth --params-n 512 --batch 1184 --layers 20 --steps 1 --warmup 1 --wiring allgather > gpurun_out/bench_r02_8gpu_depth_allgather.json 2> gpurun_out/bench_r02_8gpu_depth_allgather.err; cat gpurun_out/bench_r02_8gpu_depth_allgather.json
$TR bench.py --gpus 8 --workload depth --params-n 512 --batch 1184 --layers 20 --steps 1 --warmup 1 --wiring local > gpurun_out/bench_r02_8gpu_depth_local.json 2> gpurun_out/bench_r02_8gpu_depth_local.err; cat gpurun_out/bench_r02_8gpu_depth_local.json
tail -n 3 gpurun_out/bench_r02_8gpu_depth_*.err | cut -c1-300
