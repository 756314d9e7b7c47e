cd /root/repo
for fused in 0 1; do
TL_BENCH_FUSED=$fused timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 1 --warmup 1 --no-e2e 2>gpurun_out/b.err | tail -1 > gpurun_out/n8_fused$fused.json
python -c "
import json
d=json.loads(open('gpurun_out/n8_fused$fused.json').read())
print('N=8 fused=$fused value %.4e ms/step %.1f iters %s' % (d['value'], d['ms_per_step'], d['config']['cg_iterations_per_step']))
" || tail -3 gpurun_out/b.err
done
