# Round-end measurement job (GPU box): parity tests, bench lines of the four processes, ncu launch list and full
# captures of the dominant kernel, phase timers.  Outputs under gpurun_out/; the summaries are copied to profiles/.
TAG=${1:-r01c}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_$TAG.log
for p in 1_gg_ttxgg 1_gg_ttx 1_gg_ttxg 1_gg_ttxggg; do
  python bench.py --steps 5 --warmup 3 --process $p 2> gpurun_out/bench_${TAG}_$p.err | tail -1 > gpurun_out/bench_${TAG}_$p.json
done
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_${TAG}_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_ttxgg.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:smatrix_kernel_hp -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_ttxgg_integrand python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_ttxgg.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:smatrix_kernel_hp -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_ttxggg_integrand python bench.py --steps 2 --warmup 1 --no-cpu-baseline --process 1_gg_ttxggg > gpurun_out/ncu_full_ttxggg.log 2>&1
python tools/profile_phases.py run 262144 tools/bin/libmfp_1_gg_ttxgg_prof.so > gpurun_out/phases_${TAG}_ttxgg.log 2>&1
python tools/profile_phases.py run 16384 tools/bin/libmfp_1_gg_ttxggg_prof.so > gpurun_out/phases_${TAG}_ttxggg.log 2>&1
cat gpurun_out/pytest_gpu_$TAG.log; cut -c1-160 gpurun_out/bench_${TAG}_*.json
