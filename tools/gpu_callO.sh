mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2h_bench_default_4gpu.json 2> gpurun_out/bench_4gpu.err
tail -c 600 gpurun_out/bench_4gpu.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2h_bench_default_4gpu.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['e2e']['value']), d['n_gpus'])
c=d.get('configs',{})
if c: print('   c2', c['configs[2]']['ms_per_step'], c['configs[2]']['value'], 'c3', c['configs[3]']['controller_steps_per_s'], c['configs[3]']['exchange_ms_each'], c['configs[3]']['n_gpus'])
P
