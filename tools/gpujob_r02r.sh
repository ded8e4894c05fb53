#!/bin/bash
# round 2: bank-conflict-free amplitude-buffer map chosen per process + fragment loads of 4-variant objects
cd "$(dirname "$0")/.."
C=tools/bin/libmfp_1_gg_ttxggg; B=tools/bin/libmfp_1_gg_ttxgg
bash tools/gpujob_ab.sh r02r_ttxgg_swizzle 2 262144 600 ${B}_old.so ${B}_swz.so
bash tools/gpujob_ab.sh r02r_ttxggg_swizzle 3 16384 6 ${C}_old.so ${C}_swz.so
