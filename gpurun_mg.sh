cd /root/repo
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multigpu.py -x -q -k "8 or (4 and cg)" 2>&1 | tail -3
for n in 8 4; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 1 --warmup 1 --no-e2e 2>gpurun_out/bench_n$n.err | tail -1 > gpurun_out/bench_n$n.json
python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n$n.json').read())
print('N=$n value %.4e ms/step %.1f iters %s ms/iter %.4f %s' % (d['value'], d['ms_per_step'], d['config']['cg_iterations_per_step'], d['ms_per_step']/ min(10000,(d['config']['cg_iterations_per_step'][0]+1)), d['config']['workload']))
" || tail -5 gpurun_out/bench_n$n.err
done
