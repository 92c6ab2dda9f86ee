mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_umma_gpu.py tests/test_modules_gpu.py tests/test_train_step_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo TEST FAILED; exit 1; fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_j.log 2>&1; tail -1 gpurun_out/r02_bench_j.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches_per_step'], d['roofline']['frac'])"
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k "regex:wgrad_reduce" --csv --log-file gpurun_out/r02_reduce.csv python profiles/run_step.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_reduce.csv')) if len(r)>10]
h=rows[0]; iv=h.index('Metric Value')
v=[float(r[iv].replace(',','')) for r in rows[1:]]
print('wgrad_reduce launches',len(v),'total us',sum(v)/1e3,'avg',sum(v)/1e3/len(v))
PY
