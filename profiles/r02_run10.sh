mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_step_gpu.py tests/test_modules_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_d.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_d.log | head -2; grep -o '"gpu_launches_per_step": [0-9]*' gpurun_out/r02_bench_d.log
tail -c 400 gpurun_out/r02_bench_d.log
