#!/bin/bash
# round 2, final kernels on the 8 GPUs of one box: the strong-scaling bench line (1e8 events per iteration, configs 1/2/4 inside)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 \
  2> gpurun_out/r02final_bench_8gpu.err | tail -1 > gpurun_out/r02final_bench_8gpu.json
cut -c1-400 gpurun_out/r02final_bench_8gpu.json
