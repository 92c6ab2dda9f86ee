set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_umma_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r02_pytest_umma.log
cat gpurun_out/r02_pytest_umma.log
MSMC_BENCH_DUMP=r02_shapes_b.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_b.log 2>&1
tail -c 1500 gpurun_out/r02_bench_b.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_b.log | head -2
head -40 gpurun_out/r02_shapes_b.txt
