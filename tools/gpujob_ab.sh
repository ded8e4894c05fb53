#!/bin/bash
# generic A/B job:  bash tools/gpujob_ab.sh <tag> <k> <nevents> <parity events> lib1 lib2 ... [-- prof1 prof2 ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; k=$2; nev=$3; npar=$4; shift 4
libs=(); profs=(); mode=libs
for a in "$@"; do
  if [ "$a" = "--" ]; then mode=profs; continue; fi
  if [ $mode = libs ]; then libs+=("$a"); else profs+=("$a"); fi
done
{
  python tools/time_smatrix.py $nev "${libs[@]}" "${libs[@]}"
  python tools/check_parity.py $k $npar "${libs[@]}"
  if [ ${#profs[@]} -gt 0 ]; then python tools/profile_phases.py run $nev "${profs[@]}"; fi
} 2>&1 | tee gpurun_out/${tag}.log
