#!/bin/bash
# round-2 call R (2 GPUs): final state -- whole GPU suite, kernel roofline table, stamps and bench lines at N = 1 and 2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2r_gputests.log
tail -5 gpurun_out/r2r_gputests.log
timeout 300 python tools/kernel_roofline.py > gpurun_out/r2r_kernels_roofline.txt 2>&1
cat gpurun_out/r2r_kernels_roofline.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
S=gpurun_out/r2r_stamps.txt
timeout 300 python tools/stamps.py --tag n1_fused > $S 2>&1
timeout 300 python tools/stamps.py --tag n1_three --fused 0 >> $S 2>&1
timeout 300 $TR --master-port 29591 tools/stamps.py --tag n2_fused >> $S 2>&1
grep "^#\|^  [0-9]" $S
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err
timeout 600 $TR --master-port 29595 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2r_bench_n1.json", "gpurun_out/r2r_bench_n2.json"):
    d=json.load(open(f))
    print("N=%d value %.4e e2e %.4e ms/iter %.4f  parity %s frac %.3f in_loop %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]*d["steps"]/sum(d["config"]["cg_iterations_per_step"]), d["parity"]["n_chunk_bit_exact"], d["roofline"]["frac"], d["roofline"]["in_loop"]))
    print({k: round(v["ms"],4) for k,v in d["roofline"]["kernels"].items()}, d["clocks"], d["cpu_baseline"])
    print(d["iteration_profile"])
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2r_bench_ref_n1.json 2>> gpurun_out/r2r_bench_n1.err; head -c 600 gpurun_out/r2r_bench_ref_n1.json
