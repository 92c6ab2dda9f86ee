mkdir -p gpurun_out
export BENCH_REUSE_CONFIGS=persist-nacc1
for dry in 0 1 2 3 4 8 7 15; do echo "== DRY=$dry"; MSMC_PERSIST_DRY=$dry timeout 600 python profiles/bench_reuse.py mrf32k3 mrf32k11 mrf64k11 mrf128k11 ffn2 2>&1 | cut -c1-100; done > gpurun_out/r02_bench_reuse_dry.txt
cat gpurun_out/r02_bench_reuse_dry.txt
