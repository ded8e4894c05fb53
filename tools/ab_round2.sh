#!/bin/bash
# First GPU job of the next round: A/B of the host-verified, not yet measured kernel flavours (plan).
#   bash tools/ab_round2.sh build      here (CPU): builds the variants into tools/bin/ (they travel with the snapshot)
#   gpurun --timeout 300 -- 'bash tools/ab_round2.sh run'      on the GPU box: timing + parity of every variant
set -e
cd "$(dirname "$0")/.."
if [ "${1:-build}" = "build" ]; then
  rm -f tools/bin/*
  python tools/build_variants.py 2 base: chain:CHAIN=1 chain_e1:CHAIN=1,E=1,MINBLOCKS=4 chain_mt4:CHAIN=1,MT=4 e1:E=1,MINBLOCKS=4
  python tools/build_variants.py 3 base: chain:CHAIN=1 chain_mt2:CHAIN=1,MT=2
  rm -f tools/bin/var_*
  ls -la tools/bin
else
  mkdir -p gpurun_out
  {
    python tools/time_smatrix.py 262144 tools/bin/libmfp_1_gg_ttxgg_base.so tools/bin/libmfp_1_gg_ttxgg_chain.so \
      tools/bin/libmfp_1_gg_ttxgg_chain_e1.so tools/bin/libmfp_1_gg_ttxgg_chain_mt4.so tools/bin/libmfp_1_gg_ttxgg_e1.so \
      tools/bin/libmfp_1_gg_ttxgg_base.so tools/bin/libmfp_1_gg_ttxgg_chain.so
    python tools/time_smatrix.py 16384 tools/bin/libmfp_1_gg_ttxggg_base.so tools/bin/libmfp_1_gg_ttxggg_chain.so \
      tools/bin/libmfp_1_gg_ttxggg_chain_mt2.so tools/bin/libmfp_1_gg_ttxggg_base.so tools/bin/libmfp_1_gg_ttxggg_chain.so
    python tools/check_parity.py 2 600 tools/bin/libmfp_1_gg_ttxgg_chain.so tools/bin/libmfp_1_gg_ttxgg_chain_e1.so
    python tools/check_parity.py 3 6 tools/bin/libmfp_1_gg_ttxggg_chain.so
  } 2>&1 | tee gpurun_out/ab_round2.log
fi
