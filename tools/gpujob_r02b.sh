#!/bin/bash
# round 2: flat unit records + prefetch; NE events per thread in the unit phases (g g > t t~ g g)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=tools/bin/libmfp_1_gg_ttxgg
{
  python tools/time_smatrix.py 262144 ${B}_r11.so ${B}_r21.so ${B}_r22.so ${B}_r12.so ${B}_r21mt4.so ${B}_r11.so ${B}_r21.so
  python tools/check_parity.py 2 600 ${B}_r11.so ${B}_r21.so
  python tools/profile_phases.py run 262144 ${B}_r11p.so ${B}_r21p.so
} 2>&1 | tee gpurun_out/r02b_ab.log
