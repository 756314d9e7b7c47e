#!/bin/bash
# round-2 call Q (4 GPUs): Chebyshev / PPCG to convergence (device-side switch rule, in-kernel halos, folded norms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_bit_exact_gpu.py tests/test_dropin_gpu.py -q 2>&1 | tail -6 > gpurun_out/r2q_tests.log
timeout 900 python -m pytest tests/test_multigpu.py -q -k "cheby or ppcg" 2>&1 | tail -6 >> gpurun_out/r2q_tests.log
cat gpurun_out/r2q_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
for s in cheby ppcg; do
  timeout 600 python bench.py --steps 1 --warmup 1 --solver $s --mesh 4000 4000 --max-iters 30000 --no-cpu --no-parity --no-e2e > gpurun_out/r2q_bench_n1_$s.json 2> gpurun_out/r2q_bench.err
  timeout 600 $TR --master-port 29581 bench.py --gpus 4 --steps 1 --warmup 1 --solver $s --mesh 8000 8000 --max-iters 30000 --no-extra --no-parity --no-e2e > gpurun_out/r2q_bench_n4_$s.json 2>> gpurun_out/r2q_bench.err
done
python - <<'PY'
import json
for f in ("n1_cheby","n4_cheby","n1_ppcg","n4_ppcg"):
    try:
        d=json.load(open("gpurun_out/r2q_bench_%s.json" % f))
        print(f, "value %.4e ms/step %.1f iters/step %s launches %d temp %.12f" % (d["value"], d["ms_per_step"], d["config"]["cg_iterations_per_step"], d["gpu_launches"], d["summary"]["temp"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -5 gpurun_out/r2q_bench.err
