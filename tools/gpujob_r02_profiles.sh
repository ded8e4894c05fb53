#!/bin/bash
# round 2: ncu --set full captures of the dominant kernel of the four BASELINE processes inside bench.py's integrand,
# phase timers, and the bench lines themselves.  Outputs under gpurun_out/ (summaries are copied to profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none -f"
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs"
$NCU -k regex:smatrix_kernel_hp -s 3 -c 1 -o gpurun_out/r02_prof_ttxgg_integrand $B --events 8388608 > gpurun_out/ncu_ttxgg.log 2>&1
$NCU -k regex:smatrix_kernel_hp -s 3 -c 1 -o gpurun_out/r02_prof_ttxggg_integrand $B --process 1_gg_ttxggg --events 262144 > gpurun_out/ncu_ttxggg.log 2>&1
$NCU -k regex:smatrix_kernel_hp -s 3 -c 1 -o gpurun_out/r02_prof_ttxg_integrand $B --process 1_gg_ttxg --events 8388608 > gpurun_out/ncu_ttxg.log 2>&1
$NCU -k regex:integrand_kernel -s 3 -c 1 -o gpurun_out/r02_prof_ttx_integrand $B --process 1_gg_ttx --events 33554432 > gpurun_out/ncu_ttx.log 2>&1
ls -la gpurun_out/*.ncu-rep
