#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for pre in -1 0 1 2 3; do
  TL_PDL=1 TL_PW=1:3:0 TL_PW_PRE=$pre timeout 300 python tools/bm5_check.py 3 2>&1 | sed "s/^/PRE=$pre /" >> gpurun_out/r2f_bm5.txt
done
done
cat gpurun_out/r2f_bm5.txt
SPECS="0:4:0:0:2 1:3:0:0:2 1:4:0:0:2 0:4:0:0:0"
for cfg in "0 100 3" "1 100 -1" "1 100 3" "0 -1 3" "1 -1 -1" "1 -1 3" "1 50 3" "0 50 3"; do
  set -- $cfg
  TL_PDL=$1 TL_PW_CARVEOUT=$2 TL_PW_PRE=$3 timeout 300 python tools/loop_rate.py 4000 $SPECS 2>&1 | sed "s/^/CARVE=$2 PRE=$3 /" >> gpurun_out/r2f_loop.txt
done
cat gpurun_out/r2f_loop.txt
