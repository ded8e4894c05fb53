#!/bin/bash
# round 2, first GPU job: the colour-reduced plan against the diagram list (g g > t t~ g g): timing, parity, phases
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
  python tools/time_smatrix.py 262144 tools/bin/libmfp_1_gg_ttxgg_red.so tools/bin/libmfp_1_gg_ttxgg_legacy.so tools/bin/libmfp_1_gg_ttxgg_red.so tools/bin/libmfp_1_gg_ttxgg_legacy.so
  python tools/check_parity.py 2 600 tools/bin/libmfp_1_gg_ttxgg_red.so tools/bin/libmfp_1_gg_ttxgg_legacy.so
  python tools/profile_phases.py run 262144 tools/bin/libmfp_1_gg_ttxgg_red_prof.so tools/bin/libmfp_1_gg_ttxgg_legacy_prof.so
} 2>&1 | tee gpurun_out/r02a_ab.log
