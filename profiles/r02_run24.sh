mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_modules_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['ms_per_step'], d['kernel_families']['msmc_conv_wgrad_umma'], d['roofline']['frac'])"
