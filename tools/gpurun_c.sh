#!/bin/bash
# round-2 call C (1 GPU): ncu --set full of the resident CG loop, bulk pipeline vs register kernel
mkdir -p gpurun_out
TL_PW=1:3:0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_calc -s 24 -c 4 -o gpurun_out/r2c_bulk -f python tools/profile_cg.py --iters 20 > gpurun_out/r2c_bulk.log 2>&1
TL_PW=0:4:0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cg_calc -s 24 -c 2 -o gpurun_out/r2c_reg -f python tools/profile_cg.py --iters 20 > gpurun_out/r2c_reg.log 2>&1
tail -3 gpurun_out/r2c_bulk.log gpurun_out/r2c_reg.log
ls -la gpurun_out/*.ncu-rep
