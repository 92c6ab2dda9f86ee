mkdir -p gpurun_out
timeout 600 python profiles/trace_step.py r02a 2>&1 | tail -3
ls -la gpurun_out | tail -3
