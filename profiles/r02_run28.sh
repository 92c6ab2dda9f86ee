mkdir -p gpurun_out
for n in 960 3840; do
timeout 200 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,launch__grid_size --clock-control none --kernel-name-base demangled -k "regex:vq_search" --csv --log-file gpurun_out/r02_vq_small_$n.csv python profiles/run_vq_case.py $n 256 umma > /dev/null 2>&1
grep -E "gpu__time_duration|sm__cycles_elapsed" gpurun_out/r02_vq_small_$n.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | paste - - | head -10
done
