#!/bin/bash
# compute-sanitizer on the packed-unit kernels (g g > t t~ g g: packed units; g g > t t~ g g g: split units with warp
# shuffles, JAMPs in Tensor Memory, inputs fetched ahead) and on the integrand pipeline; small samples (the tools are slow)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=madflow_b200/lib
{
  for tool in memcheck racecheck synccheck; do
    echo "=== compute-sanitizer --tool $tool  mfp_smatrix  libmfp_1_gg_ttxgg.so (600 events)"
    compute-sanitizer --tool $tool --print-limit 5 python tools/time_smatrix.py 600 $L/libmfp_1_gg_ttxgg.so 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|ev/s" | head -12
    echo "=== compute-sanitizer --tool $tool  mfp_smatrix  libmfp_1_gg_ttxggg.so (300 events)"
    compute-sanitizer --tool $tool --print-limit 5 python tools/time_smatrix.py 300 $L/libmfp_1_gg_ttxggg.so 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|ev/s" | head -12
  done
  echo "=== compute-sanitizer --tool memcheck  fused integrand (segmented mode) + deterministic accumulation (pytest -k fused_integrand...ttxgg)"
  compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests -m gpu -x -q -k "fused_integrand_generated_processes and ttxgg-2 and hp" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Error" | head
} 2>&1 | tee gpurun_out/r02final_compute_sanitizer.txt
