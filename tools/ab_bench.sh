#!/bin/bash
# tools/ab_bench.sh [configs...] — A/B the in-tree library against kernel variants built into build_variants/*.so (LMPC_B200_SO
# override): one line per (variant, config) with the step time and the per-kernel times of configs[2].  Run on the GPU box.
CFGS=${@:-1 2}
for so in racinglmpc_b200/liblmpc_b200.so build_variants/*.so; do
  [ -f "$so" ] || continue
  for cfg in $CFGS; do
    LMPC_B200_SO=$PWD/$so python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d.get('kernels',{}); print('$so', 'config', $cfg, 'ms', round(d['ms_per_step'],4), 'solved', d['config']['solved_fraction'], 'iters', round(d['config']['ipm_iters_mean'],3), {n.split('_kernel')[0]: round(v['ms'],4) for n,v in k.items()})"
  done
done
