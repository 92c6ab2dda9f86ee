mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_ops_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5
export BENCH_REUSE_CONFIGS=one-tile,persist-auto
timeout 600 python profiles/bench_reuse.py > gpurun_out/r02_bench_reuse_d.txt 2>&1
cat gpurun_out/r02_bench_reuse_d.txt
for tg in 1 2; do echo "== TG=$tg"; MSMC_PERSIST_TG=$tg BENCH_REUSE_CONFIGS=persist-auto timeout 600 python profiles/bench_reuse.py mrf32k3 mrf32k11 mrf64k3 mrf64k11 2>&1 | cut -c1-100; done
MSMC_BENCH_DUMP=r02_shapes_c.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_c.log | head -2
