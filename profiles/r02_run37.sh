mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_ops_gpu.py tests/test_modules_gpu.py tests/test_train_step_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo TEST FAILED; exit 1; fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_l.log 2>&1; tail -1 gpurun_out/r02_bench_l.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches_per_step'], d['roofline']['frac'], d['roofline']['kernel'], d['roofline']['avg_launch_us']); print({k:v['ms_per_step'] for k,v in list(d['kernel_families'].items())[:6]})"
