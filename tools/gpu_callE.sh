set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo pytest rc $?
tail -5 gpurun_out/pytest_gpu.log
for ns in 2 3 4; do
  LMPC_B200_E2E_SLOTS=$ns timeout 300 python bench.py --headline-only --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_e2e_slots${ns}_20.json 2> gpurun_out/bench_e2e_slots$ns.err
  LMPC_B200_E2E_SLOTS=$ns timeout 300 python bench.py --headline-only --no-cpu-baseline --steps 200 --warmup 3 > gpurun_out/bench_e2e_slots${ns}_200.json 2>> gpurun_out/bench_e2e_slots$ns.err
done
for ch in 2 8; do
  LMPC_B200_CHUNKS=$ch timeout 300 python bench.py --headline-only --no-cpu-baseline --steps 200 --warmup 3 > gpurun_out/bench_e2e_chunks${ch}_200.json 2> gpurun_out/bench_e2e_chunks$ch.err
done
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_e2e_*.json'))+['gpurun_out/bench_default.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']))
    except Exception as e: print(f, 'ERR', e)
P
cat gpurun_out/bench_default.time
