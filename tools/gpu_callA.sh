set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | head -30 > gpurun_out/lscpu.txt 2>&1
timeout 300 python tools/numa_probe.py > gpurun_out/numa_probe.json 2> gpurun_out/numa_probe.err
for mode in auto off; do
  LMPC_B200_NUMA=$mode timeout 300 python bench.py --headline-only --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_headline_numa_$mode.json 2> gpurun_out/bench_headline_numa_$mode.err
  LMPC_B200_NUMA=$mode timeout 300 python bench.py --headline-only --no-cpu-baseline --steps 200 --warmup 3 > gpurun_out/bench_headline200_numa_$mode.json 2>> gpurun_out/bench_headline_numa_$mode.err
done
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo pytest rc $?
tail -5 gpurun_out/pytest_gpu.log
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_headline*_numa_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), d['config'].get('host_numa_node'))
    except Exception as e: print(f, 'ERR', e)
P
cat gpurun_out/numa_probe.json | head -60
