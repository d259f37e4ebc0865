mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo pytest rc $?
tail -3 gpurun_out/pytest_gpu.log
bash tools/ab_bench.sh 1 2 2>&1 | tee gpurun_out/ab_endgame.txt
for so in racinglmpc_b200/liblmpc_b200.so build_variants/noendgame.so; do
  LMPC_B200_SO=$PWD/$so timeout 200 python bench.py --config 3 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); c=d['detail']; print('$so', 'c3 steps/s', round(d['value']), {k:c[k] for k in ('closed_loop_steps','ipm_iters_mean_sampled','ipm_iters_max_sampled','instances_with_flags_rank0','unsolved_steps_rank0','late_accepts_rank0')}, [round(l['mean'],2) for l in c['lap_stats_rank0']])"
done 2>&1 | tee gpurun_out/c3_endgame.txt
