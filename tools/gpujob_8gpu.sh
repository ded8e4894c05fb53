# C4 / C5 of BASELINE.json on the 8 GPUs of one box
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --process 1_gg_ttxggg 2>/dev/null | tail -1 > gpurun_out/bench_r01b_8gpu_1_gg_ttxggg.json
$TR --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r01b_8gpu_1_gg_ttxgg.json
$TR --master-port 29523 -m madflow_b200.scripts.madflow_exec --madgraph_process "g g > t t~ g g" --no_pdf -c --dr_cut -i 12 -f 6 \
   --events_per_iteration 100000000 --histograms --unweighted_events 200000 -o gpurun_out/c5_ttxgg_8gpu 2>&1 | grep "madflow" | grep -v "^\[INFO\] (madflow)" | sort | uniq | tail -30 > gpurun_out/c5_ttxgg_8gpu.log
cut -c1-200 gpurun_out/bench_r01b_8gpu_*.json; tail -12 gpurun_out/c5_ttxgg_8gpu.log
