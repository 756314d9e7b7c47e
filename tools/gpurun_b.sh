#!/bin/bash
# round-2 call B (1 GPU): bulk-pipeline p+w kernel: correctness, then sweep
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_solver_gpu.py tests/test_bit_exact_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
timeout 900 python tools/tune_pw.py > gpurun_out/r2b_tune_pw.txt 2>&1
head -20 gpurun_out/r2b_tune_pw.txt
grep -E "BEST|mode=0" gpurun_out/r2b_tune_pw.txt
sort -t'p' -k4 gpurun_out/r2b_tune_pw.txt | grep "mode=1" | sort -k11 -n | head -12
