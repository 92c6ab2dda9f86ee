mkdir -p gpurun_out
timeout 600 python profiles/trace_step.py r02g 2>&1 | tail -1
gzip -f gpurun_out/r02g_timeline.csv
