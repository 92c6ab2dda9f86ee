mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:vq_search_umma" --launch-skip 2 --launch-count 1 -f -o /tmp/r02_vqumma python profiles/run_vq_case.py 262144 256 umma > gpurun_out/ncu_vqumma.log 2>&1
tail -2 gpurun_out/ncu_vqumma.log
ncu -i /tmp/r02_vqumma.ncu-rep --page raw --csv > gpurun_out/r02_vqumma_raw.csv 2>/dev/null
ncu -i /tmp/r02_vqumma.ncu-rep --page details > gpurun_out/r02_vqumma_details.txt 2>/dev/null
ncu -i /tmp/r02_vqumma.ncu-rep --page source --csv --print-kernel-base demangled > gpurun_out/r02_vqumma_source.csv 2>/dev/null
