#!/bin/bash
# round-2 call L (2 GPUs): in-loop sweep of the calc_pw tile height, 1 and 2 ranks
mkdir -p gpurun_out
SPECS="0:1 48:1 52:1 54:1 56:1 58:1 60:1 62:1 63:1 64:1 66:1 68:1 72:1 80:1 28:1 32:1 64:2 48:2 32:2"
timeout 900 python tools/tune_rows_loop.py $SPECS > gpurun_out/r2l_rows_n1.txt 2>&1
cat gpurun_out/r2l_rows_n1.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/tune_rows_loop.py $SPECS > gpurun_out/r2l_rows_n2.txt 2>&1
grep "^ranks" gpurun_out/r2l_rows_n2.txt
