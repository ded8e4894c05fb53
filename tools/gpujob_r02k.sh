#!/bin/bash
# round 2: tuning around the TMEM-parked JAMPs
cd "$(dirname "$0")/.."
C=tools/bin/libmfp_1_gg_ttxggg; B=tools/bin/libmfp_1_gg_ttxgg
bash tools/gpujob_ab.sh r02k_ttxggg_tmem_tuning 3 16384 6 ${C}_tm.so ${C}_tm_ts0.so ${C}_tm_ts2.so ${C}_tm_free.so ${C}_tm_ncg4.so
python tools/time_smatrix.py 262144 ${B}_mb1.so ${B}_mb1tm.so ${B}_mb1.so ${B}_mb1tm.so 2>&1 | tee gpurun_out/r02k_ttxgg_one_block_per_sm_tmem.log
