#!/bin/bash
# round-2 call M (4 GPUs): multi-GPU tests at world 2 and 4 (left/right neighbours), stamps and bench at N = 4
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2m_gpus.txt
timeout 1500 python -m pytest tests/test_multigpu.py -q 2>&1 | tail -30 > gpurun_out/r2m_multigpu_tests.log
tail -6 gpurun_out/r2m_multigpu_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
S=gpurun_out/r2m_stamps.txt
timeout 300 $TR --master-port 29561 tools/stamps.py --tag n4_fused --fused 2 > $S 2>&1
timeout 300 $TR --master-port 29562 tools/stamps.py --tag n4_three --fused 0 >> $S 2>&1
grep "^#\|^  [0-9]" $S
timeout 900 $TR --master-port 29565 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2m_bench_n4.json 2> gpurun_out/r2m_bench_n4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2m_bench_n4.json"))
print("N=%d value %.4e e2e %.4e ms/iter %.4f  parity %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]*d["steps"]/sum(d["config"]["cg_iterations_per_step"]), d["parity"]))
print(d["loop_form_tuning"]); print(d["iteration_profile"]); print(d["extra"]); print(d["config"]["iteration"])
PY
tail -3 gpurun_out/r2m_bench_n4.err
