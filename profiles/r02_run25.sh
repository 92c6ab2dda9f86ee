mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vq_umma_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8
timeout 600 python profiles/bench_vq.py > gpurun_out/r02_bench_vq_c.txt 2>&1; cat gpurun_out/r02_bench_vq_c.txt
