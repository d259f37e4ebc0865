# Final evidence capture of a round on one gpurun B200 box: GPU tests, smoke, both bench arms, probes, ncu capture + launch list, sanitizer.
# Usage: gpurun --timeout 2400 -- "bash tools/gpu_validate.sh"; results land in gpurun_out/ (copy what is to be judged into profiles/).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo pytest rc $?
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/r2h_bench_reference_arm.json 2> gpurun_out/bench_reference.err
timeout 900 python bench.py > gpurun_out/r2h_bench_default_1gpu.json 2> gpurun_out/bench_default.err
timeout 300 python tools/e2e_probe.py > gpurun_out/r2h_e2e_probe.txt 2>&1
timeout 300 python benchmarks/horizon_sweep.py 2>/dev/null | tail -4 > gpurun_out/r2h_horizon_sweep.jsonl
timeout 300 python benchmarks/rollout_mc.py --batch 4096 --mode independent 2>/dev/null | tail -1 > gpurun_out/r2h_rollout_mc_1gpu_independent.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ftocp_kernel -c 1 -s 5 -o gpurun_out/r2h_ftocp_N12_M0 -f python bench.py --headline-only --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches_bench_default_steps3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -2 gpurun_out/sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/sanitizer_racecheck.log 2>&1; tail -2 gpurun_out/sanitizer_racecheck.log
timeout 400 compute-sanitizer --tool synccheck python tools/sanitize.py > gpurun_out/sanitizer_synccheck.log 2>&1; tail -2 gpurun_out/sanitizer_synccheck.log
python - <<'P'
import json
for f in ['gpurun_out/r2h_bench_default_1gpu.json','gpurun_out/r2h_bench_reference_arm.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), d.get('cpu_baseline',{}).get('value'))
    c=d.get('configs',{})
    if c: print('   c2', c['configs[2]']['ms_per_step'], c['configs[2]']['e2e']['value'], 'c3', c['configs[3]']['controller_steps_per_s'])
P
tail -5 gpurun_out/r2h_e2e_probe.txt
