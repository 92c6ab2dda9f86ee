mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vq_umma_gpu.py tests/test_ops_gpu.py -k vq -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo TEST FAILED; exit 1; fi
timeout 300 python profiles/bench_vq.py > gpurun_out/r02_bench_vq_e.txt 2>&1; cat gpurun_out/r02_bench_vq_e.txt | tail -16
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_train_step_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_g.log 2>&1; tail -c 1500 gpurun_out/r02_bench_g.log
