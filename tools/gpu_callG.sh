mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo pytest rc $?
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for i in 1 2; do
timeout 900 python bench.py > gpurun_out/bench_default_$i.json 2> gpurun_out/bench_default_$i.err
done
timeout 300 python bench.py --headline-only --no-cpu-baseline --steps 200 --warmup 3 > gpurun_out/bench_headline200.json 2>/dev/null
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_default_*.json'))+['gpurun_out/bench_headline200.json','gpurun_out/bench_reference.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), d.get('cpu_baseline',{}).get('value'), d['config'].get('host_numa_node'))
        c=d.get('configs',{})
        if c: print('   c2', c['configs[2]']['ms_per_step'], c['configs[2]']['e2e']['value'], 'c3', c['configs[3]']['controller_steps_per_s'])
    except Exception as e: print(f, 'ERR', e)
P
