mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_modules_gpu.py tests/test_train_step_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo TEST FAILED; exit 1; fi
for f in 0 1; do MSMC_ATTN_BWD_FORK=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('fork=$f', d['ms_per_step'], d['e2e']['ms_per_step'])"; done
