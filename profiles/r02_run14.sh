mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -8
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_modules_gpu.py -m gpu -x -q -p no:cacheprovider -k "predictor" 2>&1 | tail -12
timeout 900 python bench.py --config am --steps 10 --warmup 3 > gpurun_out/r02_bench_am_a.log 2>&1
tail -c 2500 gpurun_out/r02_bench_am_a.log
