cd /root/repo
for tune in "1:8:4,2:8:4" "1:16:4,2:16:4" "1:32:4,2:32:4" "1:16:2,2:16:4"; do
for fused in 1 2; do
TL_TUNE=$tune TL_BENCH_FUSED=$fused timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --no-e2e --max-iters 3000 2>gpurun_out/b.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=2 tune=$tune fused=$fused value %.4e ms/iter %.4f' % (d['value'], d['ms_per_step']/ 3000))
" || tail -5 gpurun_out/b.err
done
done
