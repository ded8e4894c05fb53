#!/bin/bash
# round 2, after the last kernel change (inputs of the next event fetched ahead, g g > t t~ g g g): the bench line with
# configs 1/2/4 inside, the ncu capture of g g > t t~ g g g, phase timers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02final
python bench.py 2> gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench_1gpu.json
NCU="ncu --set full --import-source on --clock-control none -f"
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs"
$NCU -k regex:smatrix_kernel_hp -s 3 -c 1 -o gpurun_out/${T}_prof_ttxggg_integrand $B --process 1_gg_ttxggg --events 262144 > gpurun_out/${T}_ncu_ttxggg.log 2>&1
python tools/profile_phases.py run 16384 tools/bin/libmfp_1_gg_ttxggg_prof.so > gpurun_out/${T}_phases_ttxggg.txt 2>&1
cut -c1-300 gpurun_out/${T}_bench_1gpu.json; cat gpurun_out/${T}_phases_ttxggg.txt
