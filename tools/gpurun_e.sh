#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for cfg in "0 0:4:0" "1 0:4:0" "0 1:3:0" "1 1:3:0" "1 1:4:0"; do
  set -- $cfg
  TL_PDL=$1 TL_PW=$2 timeout 300 python tools/bm5_check.py 4 >> gpurun_out/r2e_bm5.txt 2>&1
done
done
cat gpurun_out/r2e_bm5.txt
