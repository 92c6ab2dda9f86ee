mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_ops_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
for w in 0 1; do MSMC_WGRAD_REUSE_BN128=$w timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('BN128=$w', d['ms_per_step'], d['kernel_families']['msmc_conv_wgrad_umma'])"; done
