mkdir -p gpurun_out
MSMC_BENCH_DUMP=r02_shapes_d.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_f.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02_bench_f.log') if x.startswith('{')][-1]
d=json.loads(l)
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches_per_step'])
print(json.dumps(d['roofline'],indent=0)[:900])
for k,v in list(d['kernel_families'].items())[:12]: print(k,v)
PY
head -30 gpurun_out/r02_shapes_d.txt
