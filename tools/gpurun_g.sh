#!/bin/bash
# round-2 call G (2 GPUs): tail-gather + LL multi-rank protocol: all GPU tests, stamps at N=1/2, bench N=1/2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g_gpus.txt
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2g_gputests.log
tail -5 gpurun_out/r2g_gputests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python tools/stamps.py --tag n1_fused > gpurun_out/r2g_stamps.txt 2>&1
timeout 300 python tools/stamps.py --tag n1_three --fused 0 >> gpurun_out/r2g_stamps.txt 2>&1
timeout 300 $TR --master-port 29521 tools/stamps.py --tag n2_fused --fused 2 >> gpurun_out/r2g_stamps.txt 2>&1
timeout 300 $TR --master-port 29522 tools/stamps.py --tag n2_three --fused 0 >> gpurun_out/r2g_stamps.txt 2>&1
TL_PDL=0 timeout 300 $TR --master-port 29523 tools/stamps.py --tag n2_fused_nopdl --fused 2 >> gpurun_out/r2g_stamps.txt 2>&1
grep -v "^W\|^\*\*\*\|^$" gpurun_out/r2g_stamps.txt | tail -40
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
timeout 600 $TR --master-port 29524 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
head -c 1200 gpurun_out/r2g_bench_n1.json; echo
head -c 3000 gpurun_out/r2g_bench_n2.json; echo
tail -5 gpurun_out/r2g_bench_n2.err
