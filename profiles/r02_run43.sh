mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_vq_umma_gpu.py tests/test_ops_gpu.py -k vq -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo TEST FAILED; exit 1; fi
timeout 300 python profiles/bench_vq.py > gpurun_out/r02_bench_vq_g.txt 2>&1; cat gpurun_out/r02_bench_vq_g.txt | tail -14
timeout 200 python profiles/vq_small_probe.py 2>&1 | tail -8 | cut -c1-60
