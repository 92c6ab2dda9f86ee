timeout 900 python -m pytest tests/test_umma_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
timeout 900 python -m pytest tests/test_train_step_gpu.py -m gpu -x -q -p no:cacheprovider -k "predictor" 2>&1 | grep -E "^E|passed|failed" | head
