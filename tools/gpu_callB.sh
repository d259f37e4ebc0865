set -x
mkdir -p gpurun_out
timeout 300 python tools/numa_probe.py > gpurun_out/numa_probe2.json 2> gpurun_out/numa_probe2.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo pytest rc $?
tail -5 gpurun_out/pytest_gpu.log
for sp in 1 2 3 4; do
  LMPC_B200_STEP_SPLIT=$sp timeout 300 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_split$sp.json 2> gpurun_out/bench_c2_split$sp.err
done
for sp in 1 2 4; do
  LMPC_B200_STEP_SPLIT=$sp timeout 300 python bench.py --config 3 > gpurun_out/bench_c3_split$sp.json 2> gpurun_out/bench_c3_split$sp.err
done
for mode in auto off; do
  LMPC_B200_NUMA=$mode timeout 300 python bench.py --headline-only --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_headline_numa_$mode.json 2> gpurun_out/bench_headline_numa_$mode.err
  LMPC_B200_NUMA=$mode timeout 300 python bench.py --headline-only --no-cpu-baseline --steps 200 --warmup 3 > gpurun_out/bench_headline200_numa_$mode.json 2>> gpurun_out/bench_headline_numa_$mode.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c2_split*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['config'].get('solved_fraction'))
    except Exception as e: print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/bench_c3_split*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, {k:d.get(k) for k in ('controller_steps_per_s','ms_total','kernel_launches_rank0','instances_with_flags_rank0','unsolved_steps_rank0','late_accepts_rank0')}, [l['mean'] for l in d.get('lap_stats_rank0',[])])
    except Exception as e: print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/bench_headline*_numa_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), d['config'].get('host_numa_node'))
    except Exception as e: print(f, 'ERR', e)
d=json.load(open('gpurun_out/numa_probe2.json'))
for k,v in d['h2d_d2h_gbs'].items(): print(k, v if isinstance(v,str) else (v.get('node'), v['pages'], [round(x,1) for x in v['bw']]))
P
tail -3 gpurun_out/numa_probe2.err
