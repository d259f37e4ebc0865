mkdir -p gpurun_out
for rep in 1 2 3; do for ch in auto 1 2 4; do
  if [ $ch = auto ]; then unset LMPC_B200_CHUNKS; else export LMPC_B200_CHUNKS=$ch; fi
  timeout 300 python bench.py --headline-only --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rep $rep chunks $ch steps 20 value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done; done
unset LMPC_B200_CHUNKS
bash tools/ab_bench.sh 1
