mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo pytest rc $?
tail -3 gpurun_out/pytest_gpu.log
bash tools/ab_bench.sh 1 2>&1 | tee gpurun_out/ab_gamma.txt
bash tools/ab_bench.sh 1 2>&1 | tee -a gpurun_out/ab_gamma.txt
for so in racinglmpc_b200/liblmpc_b200.so build_variants/gamma_1e-2.so build_variants/gamma_1e-3.so; do
  echo $so; LMPC_B200_SO=$PWD/$so timeout 300 python benchmarks/horizon_sweep.py 2>/dev/null | tail -4 | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['N'], round(d['solves_per_s']), d['ipm_iters_mean'], d['ipm_iters_max'], d['solved_fraction'], d['max_resid'])"
done 2>&1 | tee gpurun_out/sweep_gamma.txt
