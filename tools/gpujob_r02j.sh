#!/bin/bash
# round 2: JAMP accumulators parked in Tensor Memory (TMEMJ)
cd "$(dirname "$0")/.."
C=tools/bin/libmfp_1_gg_ttxggg; B=tools/bin/libmfp_1_gg_ttxgg
bash tools/gpujob_ab.sh r02j_ttxggg_tmem 3 16384 6 ${C}_base.so ${C}_tm.so
bash tools/gpujob_ab.sh r02j_ttxgg_tmem 2 262144 600 ${B}_base.so ${B}_tm.so
