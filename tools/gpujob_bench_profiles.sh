set -x
for p in 1_gg_ttxgg 1_gg_ttx 1_gg_ttxg 1_gg_ttxggg; do
  python bench.py --steps 5 --warmup 3 --process $p 2> gpurun_out/bench_r01b_$p.err | tail -1 > gpurun_out/bench_r01b_$p.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b_ttxgg.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:smatrix_kernel_hp -s 2 -c 1 -f -o gpurun_out/prof_r01b_ttxgg_integrand python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_ttxgg.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:smatrix_kernel_hp -s 2 -c 1 -f -o gpurun_out/prof_r01b_ttxggg_integrand python bench.py --steps 2 --warmup 1 --no-cpu-baseline --process 1_gg_ttxggg > gpurun_out/ncu_full_ttxggg.log 2>&1
cut -c1-300 gpurun_out/bench_r01b_*.json
