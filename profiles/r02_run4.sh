mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_umma_gpu.py -m gpu -x -q -p no:cacheprovider -k persistent 2>&1 | tail -5
timeout 600 python profiles/bench_reuse.py > gpurun_out/r02_bench_reuse_b.txt 2>&1
cat gpurun_out/r02_bench_reuse_b.txt
