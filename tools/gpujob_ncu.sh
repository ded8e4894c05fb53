#!/bin/bash
# one ncu --set full capture (with source) of the matrix-element kernel of a process library
#   bash tools/gpujob_ncu.sh <lib.so> <nevents> <tag>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:smatrix_kernel_hp -s 2 -c 1 -f -o gpurun_out/$3 \
  python tools/time_smatrix.py $2 $1 > gpurun_out/$3.log 2>&1
tail -3 gpurun_out/$3.log
ls -la gpurun_out/
