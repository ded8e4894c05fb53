#!/bin/bash
# round 2: NE events per thread on the two-level tables (g g > t t~ g g); reduced plan for g g > t t~ g g g
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=tools/bin/libmfp_1_gg_ttxgg
C=tools/bin/libmfp_1_gg_ttxggg
{
  python tools/time_smatrix.py 262144 ${B}_s11.so ${B}_s21.so ${B}_s22.so ${B}_s11.so ${B}_s21.so
  python tools/check_parity.py 2 600 ${B}_s21.so
  python tools/time_smatrix.py 16384 ${C}_red.so ${C}_legacy.so ${C}_red2.so ${C}_red.so ${C}_legacy.so
  python tools/check_parity.py 3 6 ${C}_red.so ${C}_red2.so
} 2>&1 | tee gpurun_out/r02c_ab.log
