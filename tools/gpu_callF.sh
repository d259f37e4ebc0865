mkdir -p gpurun_out
for ns in 2 3; do for ch in 1 2 3; do for st in 20 200; do
  LMPC_B200_E2E_SLOTS=$ns LMPC_B200_CHUNKS=$ch timeout 300 python bench.py --headline-only --no-cpu-baseline --steps $st --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('slots $ns chunks $ch steps $st value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done; done; done
