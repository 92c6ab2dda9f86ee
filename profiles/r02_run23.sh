mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
for gb in 24 8 64 100000; do MSMC_PF_GROUP_MB=$gb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('group MB=$gb', d['ms_per_step'], d['gpu_launches_per_step'])"; done
