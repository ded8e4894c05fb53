#!/bin/bash
# round 2: objects with many terms split over neighbouring lanes (TSPLIT), g g > t t~ g g g and g g > t t~ g g
cd "$(dirname "$0")/.."
C=tools/bin/libmfp_1_gg_ttxggg; B=tools/bin/libmfp_1_gg_ttxgg
bash tools/gpujob_ab.sh r02h_ttxggg_tsplit 3 16384 6 ${C}_ts0.so ${C}_ts1.so ${C}_ts2.so ${C}_ts3.so ${C}_ts4.so
bash tools/gpujob_ab.sh r02h_ttxgg_tsplit 2 262144 600 ${B}_ts0.so ${B}_ts1.so ${B}_ts2.so
