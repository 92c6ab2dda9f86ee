mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_modules_gpu.py tests/test_train_step_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo TEST FAILED; exit 1; fi
for cs in 32 16 8; do MSMC_UMMA_MIN_CS=$cs timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); f=d['kernel_families']; print('min_cs=$cs', d['ms_per_step'], {k:f[k]['ms_per_step'] for k in ('msmc_conv_forward_umma','msmc_conv_wgrad_umma','msmc_conv_forward','msmc_conv_wgrad')})"; done
