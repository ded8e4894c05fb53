#!/bin/bash
# round 2: two objects over the same legs per unit (shared sub-diagrams evaluated once)
cd "$(dirname "$0")/.."
C=tools/bin/libmfp_1_gg_ttxggg; B=tools/bin/libmfp_1_gg_ttxgg
bash tools/gpujob_ab.sh r02p_ttxgg_share 2 262144 600 ${B}_noshare.so ${B}_share.so -- ${B}_sharep.so
bash tools/gpujob_ab.sh r02p_ttxggg_share 3 16384 6 ${C}_noshare.so ${C}_share.so
