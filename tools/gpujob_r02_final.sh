#!/bin/bash
# round 2, final measurements of the packed-unit kernels: GPU tests, smoke, bench lines (own arm with configs 1/2/4
# inside, reference arm), ncu launch list, ncu --set full captures of g g > t t~ g g (g), phase timers.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02final
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/${T}_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${T}_bench_1gpu_reference.json
python bench.py 2> gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_ttxgg.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/${T}_ncu_launch.log 2>&1
NCU="ncu --set full --import-source on --clock-control none -f"
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-configs"
$NCU -k regex:smatrix_kernel_hp -s 3 -c 1 -o gpurun_out/${T}_prof_ttxgg_integrand $B --events 8388608 > gpurun_out/${T}_ncu_ttxgg.log 2>&1
$NCU -k regex:smatrix_kernel_hp -s 3 -c 1 -o gpurun_out/${T}_prof_ttxggg_integrand $B --process 1_gg_ttxggg --events 262144 > gpurun_out/${T}_ncu_ttxggg.log 2>&1
python tools/profile_phases.py run 262144 tools/bin/libmfp_1_gg_ttxgg_prof.so > gpurun_out/${T}_phases_ttxgg.txt 2>&1
python tools/profile_phases.py run 16384 tools/bin/libmfp_1_gg_ttxggg_prof.so > gpurun_out/${T}_phases_ttxggg.txt 2>&1
cat gpurun_out/${T}_pytest_gpu.log gpurun_out/${T}_smoke.log; cut -c1-300 gpurun_out/${T}_bench_1gpu.json gpurun_out/${T}_bench_1gpu_reference.json
cat gpurun_out/${T}_phases_ttxgg.txt gpurun_out/${T}_phases_ttxggg.txt
