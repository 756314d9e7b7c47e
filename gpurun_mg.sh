cd /root/repo
timeout 600 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -2
for sv in cheby ppcg cg; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --no-e2e --solver $sv 2>gpurun_out/b.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=2 $sv value %.4e ms/step %.1f iters %s launches %d' % (d['value'], d['ms_per_step'], d['config']['cg_iterations_per_step'], d['gpu_launches']))
" || tail -5 gpurun_out/b.err
done
