mkdir -p gpurun_out
timeout 600 python profiles/bench_reuse.py > gpurun_out/r02_bench_reuse_a.txt 2>&1
cat gpurun_out/r02_bench_reuse_a.txt
