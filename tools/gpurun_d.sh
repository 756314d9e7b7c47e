#!/bin/bash
# round-2 call D (1 GPU): programmatic dependent launch on/off, register vs bulk-pipeline p+w kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_solver_gpu.py tests/test_bit_exact_gpu.py tests/test_kernels_gpu.py -q 2>&1 | tail -15 > gpurun_out/r2d_tests.log
tail -3 gpurun_out/r2d_tests.log
SPECS="0:4:0:0:2 1:3:0:0:2 1:4:0:0:2 1:8:0:0:2 1:3:3:0:2 1:4:4:0:2 1:3:0:32:2 1:4:0:32:2 1:8:3:32:2 1:4:0:16:2 0:4:0:0:0"
for pdl in 0 1 1; do
  TL_PDL=$pdl timeout 600 python tools/loop_rate.py 4000 $SPECS >> gpurun_out/r2d_loop.txt 2>&1
done
cat gpurun_out/r2d_loop.txt
