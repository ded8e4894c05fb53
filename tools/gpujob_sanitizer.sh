#!/bin/bash
# compute-sanitizer on the helicity-parallel kernels and the integrand pipeline (small samples: the tools are slow)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=madflow_b200/lib
{
  for tool in memcheck racecheck synccheck; do
    for lib in $L/libmfp_1_gg_ttxg.so $L/libmfp_1_gg_ttxgg.so; do
      echo "=== compute-sanitizer --tool $tool  mfp_smatrix  $(basename $lib)"
      compute-sanitizer --tool $tool --print-limit 5 python tools/time_smatrix.py 600 $lib 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|ev/s" | head -12
    done
  done
  echo "=== compute-sanitizer --tool memcheck  fused integrand + deterministic accumulation (pytest -k fused_integrand...ttxgg)"
  compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests -m gpu -x -q -k "fused_integrand_generated_processes and ttxgg-2 and hp" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Error" | head
  echo "=== compute-sanitizer --tool racecheck  accumulate / histogram kernels (pytest -k bit_reproducible and ttx-0)"
  compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests -m gpu -x -q -k "event_histogram_and_unweighting" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|hazard" | head
} 2>&1 | tee gpurun_out/r02_compute_sanitizer.txt
