timeout 600 python -m pytest tests/test_modules_gpu.py -m gpu -q -p no:cacheprovider -k full_size_discriminator 2>&1 | tail -12
