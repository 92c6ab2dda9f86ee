mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_train_step_gpu.py tests/test_modules_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25
