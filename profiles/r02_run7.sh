mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_umma_gpu.py -m gpu -x -q -p no:cacheprovider -k persistent 2>&1 | tail -5
timeout 600 python profiles/bench_reuse.py > gpurun_out/r02_bench_reuse_c.txt 2>&1
cat gpurun_out/r02_bench_reuse_c.txt
export BENCH_REUSE_CONFIGS=persist-nacc1
for dry in 3 15; do echo "== DRY=$dry"; MSMC_PERSIST_DRY=$dry timeout 600 python profiles/bench_reuse.py mrf32k3 mrf32k11 mrf64k11 mrf128k11 ffn2 2>&1 | cut -c1-100; done > gpurun_out/r02_bench_reuse_dry2.txt
cat gpurun_out/r02_bench_reuse_dry2.txt
