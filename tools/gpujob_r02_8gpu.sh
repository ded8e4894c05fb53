#!/bin/bash
# round 2, the 8 GPUs of one box: BASELINE config 5 (full VEGAS integration of g g > t t~ g g with LHE / histogram
# output, 1e10 events) and the strong-scaling bench line (configs 1-4 inside)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29523 -m madflow_b200.scripts.madflow_exec --madgraph_process "g g > t t~ g g" --no_pdf -c --dr_cut -i 50 -f 40 \
   --events_per_iteration 200000000 --histograms --unweighted_events 400000 -o gpurun_out/r02_c5_ttxgg_8gpu 2>&1 \
   | grep "madflow" | grep -v "^\[INFO\] (madflow)" | sort | uniq | tail -80 > gpurun_out/r02_c5_ttxgg_8gpu.log
cp gpurun_out/r02_c5_ttxgg_8gpu/Events/*rank0/cross_err.txt gpurun_out/r02_c5_ttxgg_8gpu_cross_err.txt 2>/dev/null
cp gpurun_out/r02_c5_ttxgg_8gpu/Events/*rank0/histograms.json gpurun_out/r02_c5_ttxgg_8gpu_histograms.json 2>/dev/null
ls -la gpurun_out/r02_c5_ttxgg_8gpu/Events/*/ | head -30 >> gpurun_out/r02_c5_ttxgg_8gpu.log
rm -rf gpurun_out/r02_c5_ttxgg_8gpu
$TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_bench_8gpu.json
tail -25 gpurun_out/r02_c5_ttxgg_8gpu.log; cut -c1-400 gpurun_out/r02_bench_8gpu.json
