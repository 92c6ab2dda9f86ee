mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python profiles/run_step.py > gpurun_out/r02_ncu_list.log 2>&1
tail -2 gpurun_out/r02_ncu_list.log
python profiles/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_step_launches_summary.txt 2>&1
head -50 gpurun_out/r02_step_launches_summary.txt
gzip -f gpurun_out/r02_launches.csv
timeout 600 python profiles/trace_step.py r02b 2>&1 | tail -1
