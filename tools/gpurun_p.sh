#!/bin/bash
# round-2 call P (8 GPUs): the world = 8 multi-GPU tests, stamps and the bench line (with the other named configs) at N = 8
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2p_gpus.txt
timeout 900 python -m pytest tests/test_multigpu.py -q -k "8]" 2>&1 | tail -15 > gpurun_out/r2p_multigpu_tests_world8.log
tail -4 gpurun_out/r2p_multigpu_tests_world8.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
S=gpurun_out/r2p_stamps.txt
timeout 300 $TR --master-port 29571 tools/stamps.py --tag n8_fused --fused 2 > $S 2>&1
grep "^#\|^  [0-9]" $S
timeout 1200 $TR --master-port 29575 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2p_bench_n8.json 2> gpurun_out/r2p_bench_n8.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2p_bench_n8.json"))
print("N=%d value %.4e e2e %.4e ms/iter %.4f  parity %s %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]*d["steps"]/sum(d["config"]["cg_iterations_per_step"]), d["parity"]["n_chunk_bit_exact"], d["parity"]["halo_bit_exact"]))
print(d["loop_form_tuning"]); print(d["iteration_profile"]); print(d["clocks"])
for k, v in d["extra"].items(): print(k, v)
PY
tail -3 gpurun_out/r2p_bench_n8.err
