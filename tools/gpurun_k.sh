#!/bin/bash
# round-2 call K (2 GPUs): calc_pw without loop unrolling, calc_w with L1 neighbour loads; CLI tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2k_gputests.log
tail -8 gpurun_out/r2k_gputests.log
timeout 600 python tools/tune_stencil.py > gpurun_out/r2k_tune_stencil.txt 2>&1
grep "cg_calc" gpurun_out/r2k_tune_stencil.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
S=gpurun_out/r2k_stamps.txt
timeout 300 python tools/stamps.py --tag n1_fused > $S 2>&1
timeout 300 python tools/stamps.py --tag n1_three --fused 0 >> $S 2>&1
timeout 300 $TR --master-port 29541 tools/stamps.py --tag n2_fused --fused 2 >> $S 2>&1
timeout 300 $TR --master-port 29542 tools/stamps.py --tag n2_three --fused 0 >> $S 2>&1
grep "^#\|^  [0-9]" $S
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err
timeout 600 $TR --master-port 29545 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2k_bench_n1.json", "gpurun_out/r2k_bench_n2.json"):
    d=json.load(open(f))
    print("N=%d value %.4e e2e %.4e ms/iter %.4f  parity %s frac %.3f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]*d["steps"]/sum(d["config"]["cg_iterations_per_step"]), d["parity"]["n_chunk_bit_exact"], d["roofline"]["frac"]))
    print({k: round(v["ms"],4) for k,v in d["roofline"]["kernels"].items()}, d["clocks"])
PY
