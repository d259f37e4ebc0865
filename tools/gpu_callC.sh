set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo pytest rc $?
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --config 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_shadow.json 2> gpurun_out/bench_c2_shadow.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_c2_shadow.json').read().strip().splitlines()[-1])
print('c2 ms', d['ms_per_step'], 'value', d['value'], {k:(round(v['ms'],4), v.get('roofline',{}).get('frac')) for k,v in d['kernels'].items()})
P
timeout 300 python bench.py --config 3 > gpurun_out/bench_c3_shadow.json 2> gpurun_out/bench_c3_shadow.err
tail -c 1500 gpurun_out/bench_c3_shadow.json
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck.log
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck.log
