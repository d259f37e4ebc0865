set -x
mkdir -p gpurun_out
bash tools/ab_bench.sh 2 > gpurun_out/ab_k1.txt 2>&1
cat gpurun_out/ab_k1.txt
# ncu of K1 (in-tree variant): one launch, full set + source
timeout 600 ncu --set full --import-source on --clock-control none -k regex:knn_ltv -c 1 -s 4 -o gpurun_out/k1_async288 -f python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1.log 2>&1
tail -3 gpurun_out/ncu_k1.log
LMPC_B200_SO=$PWD/build_variants/k1_sync512.so timeout 600 ncu --set full --import-source on --clock-control none -k regex:knn_ltv -c 1 -s 4 -o gpurun_out/k1_sync512 -f python bench.py --config 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k1b.log 2>&1
tail -3 gpurun_out/ncu_k1b.log
ls -la gpurun_out/*.ncu-rep
