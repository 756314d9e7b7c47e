cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n1b.json').read())
print('value %.4e e2e %.4e ms/step %.1f iters %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['cg_iterations_per_step']))
for k,v in d['roofline']['kernels'].items(): print(k, '%.4f ms %.0f GB/s %.3f' % (v['ms'], v['gb_s'], v['frac']))
print('solve frac', d['roofline']['solve_frac_104B'])
"
