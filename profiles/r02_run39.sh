mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_vq_umma_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo TEST FAILED; exit 1; fi
timeout 300 python profiles/bench_vq.py > gpurun_out/r02_bench_vq_f.txt 2>&1; cat gpurun_out/r02_bench_vq_f.txt | tail -16
