#!/bin/bash
# round-2 call A (2 GPUs): all GPU tests incl. world=2 multi-GPU, bench N=1 and N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a_gputests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2a_bench_n2.json 2> gpurun_out/r2a_bench_n2.err
tail -5 gpurun_out/r2a_gputests.log
cat gpurun_out/r2a_bench_n1.json | head -c 1500
echo
cat gpurun_out/r2a_bench_n2.json | head -c 3000
tail -5 gpurun_out/r2a_bench_n2.err
