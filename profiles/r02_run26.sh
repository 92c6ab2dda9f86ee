mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_vq_umma_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo TEST FAILED; exit 1; fi
timeout 300 python profiles/bench_vq.py > gpurun_out/r02_bench_vq_d.txt 2>&1; cat gpurun_out/r02_bench_vq_d.txt | tail -16
for K in 64 256; do
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:vq_search_umma" --launch-skip 2 --launch-count 1 -f -o /tmp/r02_vqu$K python profiles/run_vq_case.py 262144 $K umma > gpurun_out/ncu_vqu$K.log 2>&1
tail -1 gpurun_out/ncu_vqu$K.log
ncu -i /tmp/r02_vqu$K.ncu-rep --page raw --csv > gpurun_out/r02_vqu${K}_raw.csv 2>/dev/null
ncu -i /tmp/r02_vqu$K.ncu-rep --page source --csv --print-kernel-base demangled > gpurun_out/r02_vqu${K}_source.csv 2>/dev/null
done
