#!/bin/bash
# round-2 call O (4 GPUs): halo columns forwarded group by group while the grid streams
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multigpu.py -q -k "cheby or ppcg or bit_identical or oracle_n_chunk" 2>&1 | tail -12 > gpurun_out/r2o_multigpu_tests.log
tail -4 gpurun_out/r2o_multigpu_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
S=gpurun_out/r2o_stamps.txt
timeout 300 $TR --master-port 29561 tools/stamps.py --tag n4_fused --fused 2 > $S 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 tools/stamps.py --tag n2_fused --fused 2 >> $S 2>&1
grep "^#\|^  [0-9]" $S
timeout 900 $TR --master-port 29565 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2o_bench_n4.json 2> gpurun_out/r2o_bench_n4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2o_bench_n4.json"))
print("N=%d value %.4e e2e %.4e ms/iter %.4f  parity %s %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]*d["steps"]/sum(d["config"]["cg_iterations_per_step"]), d["parity"]["n_chunk_bit_exact"], d["parity"]["halo_bit_exact"]))
print(d["loop_form_tuning"]); print(d["iteration_profile"]); print(d["extra"])
PY
tail -3 gpurun_out/r2o_bench_n4.err
for s in cheby ppcg; do timeout 300 $TR --master-port 29566 bench.py --gpus 4 --steps 2 --warmup 1 --solver $s --mesh 8000 8000 --max-iters 1000 --no-extra --no-parity --no-e2e > gpurun_out/r2o_bench_n4_$s.json 2>> gpurun_out/r2o_bench_n4.err; head -c 260 gpurun_out/r2o_bench_n4_$s.json; echo; done
