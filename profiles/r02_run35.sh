mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r02_final_bench.log 2>&1; tail -c 600 gpurun_out/r02_final_bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_ref.log 2>&1; tail -c 400 gpurun_out/r02_final_bench_ref.log
bash profiles/r02_final_profile.sh 2>&1 | tail -30
