#!/bin/bash
# round-2 call H (2 GPUs): vectorised generic kernels, one-pass cg_init, in-kernel halos of the one-pass Chebyshev / PPCG
# kernels, precise edge fences; stamps (incl. a no-send diagnostic)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2h_gputests.log
tail -8 gpurun_out/r2h_gputests.log
timeout 300 python tools/kernel_roofline.py > gpurun_out/r2h_kernels_roofline.txt 2>&1
cat gpurun_out/r2h_kernels_roofline.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
S=gpurun_out/r2h_stamps.txt
timeout 300 python tools/stamps.py --tag n1_fused > $S 2>&1
timeout 300 python tools/stamps.py --tag n1_three --fused 0 >> $S 2>&1
timeout 300 $TR --master-port 29531 tools/stamps.py --tag n2_fused --fused 2 >> $S 2>&1
timeout 300 $TR --master-port 29532 tools/stamps.py --tag n2_three --fused 0 >> $S 2>&1
# (TL_DBG_NOSEND was a timing-only diagnostic switch of this call; removed from the library afterwards)
TL_DBG_NOSEND=1 timeout 300 $TR --master-port 29533 tools/stamps.py --tag n2_fused_NOSEND_diagnostic --fused 2 >> $S 2>&1
TL_PDL=0 timeout 300 $TR --master-port 29534 tools/stamps.py --tag n2_fused_nopdl --fused 2 >> $S 2>&1
grep "^#\|^  [0-9]" $S
timeout 600 $TR --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench_n2.json"))
print("N=2 value %.4e  ms/iter %.4f  parity %s" % (d["value"], d["ms_per_step"]*d["steps"]/sum(d["config"]["cg_iterations_per_step"]), d["parity"]["n_chunk_bit_exact"]))
print({k: round(v["ms"],4) for k,v in d["roofline"]["kernels"].items()})
print(d["extra"])
PY
tail -3 gpurun_out/r2h_bench_n2.err
for s in cheby ppcg; do timeout 300 $TR --master-port 29536 bench.py --gpus 2 --steps 2 --warmup 1 --solver $s --mesh 8000 4000 --max-iters 1000 --no-extra --no-parity --no-e2e > gpurun_out/r2h_bench_n2_$s.json 2>> gpurun_out/r2h_bench_n2.err; head -c 400 gpurun_out/r2h_bench_n2_$s.json; echo; done
