# 8-GPU weak and strong scaling with the final kernel of round 2 (v18)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r02_v18_8gpu_weak.json 2> gpurun_out/bench_r02_v18_8gpu_weak.err; cut -c1-300 gpurun_out/bench_r02_v18_8gpu_weak.json
$TR bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --scaling strong --batch 4096 > gpurun_out/bench_r02_v18_8gpu_strong.json 2> gpurun_out/bench_r02_v18_8gpu_strong.err; cut -c1-300 gpurun_out/bench_r02_v18_8gpu_strong.json
