#!/bin/bash
# tools/ab_bench.sh — A/B the in-tree library against kernel variants built into build_variants/*.so (LMPC_B200_SO override):
# one JSON line per (variant, config) with ms_per_step only.  Run on the GPU box.
for so in racinglmpc_b200/liblmpc_b200.so build_variants/*.so; do
  [ -f "$so" ] || continue
  for cfg in 1 2; do
    LMPC_B200_SO=$PWD/$so python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', 'config', $cfg, 'ms', round(d['ms_per_step'],4), 'solved', d['config']['solved_fraction'], 'iters', round(d['config']['ipm_iters_mean'],3))"
  done
done
