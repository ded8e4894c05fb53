# Last GPU jobs of round 1 (12 GPU-minutes left).  Stage "new": the PDF / q q~ / multi-subprocess parity tests and the
# cost of the luminosity; stage "all": the rest of the GPU suite, smoke() and one bench line.  Every step writes its
# own file so that a clamped run still reports.
mkdir -p gpurun_out
if [ "${1:-new}" = "new" ]; then
  ( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "pdf or qqbar or multi_process" 2>&1 | tail -100 ) > gpurun_out/pytest_new_r01e.log
  tail -25 gpurun_out/pytest_new_r01e.log
  ( timeout 150 python tools/time_pdf.py 4000000 2>&1 | tail -12 ) > gpurun_out/time_pdf_r01e.log
  cat gpurun_out/time_pdf_r01e.log
else
  ( timeout 420 python -m pytest tests -m gpu -x -q --tb=short -k "not (pdf or qqbar or multi_process)" 2>&1 | tail -30 ) > gpurun_out/pytest_gpu_r01e.log
  tail -6 gpurun_out/pytest_gpu_r01e.log
  ( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/smoke_r01e.log
  cat gpurun_out/smoke_r01e.log
  timeout 240 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_r01e.err | tail -1 > gpurun_out/bench_r01e_1_gg_ttxgg.json
  cut -c1-300 gpurun_out/bench_r01e_1_gg_ttxgg.json
fi
