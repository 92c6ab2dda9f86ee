mkdir -p gpurun_out
timeout 600 python profiles/aten_sites.py > gpurun_out/r02_aten_sites.txt 2>&1; tail -50 gpurun_out/r02_aten_sites.txt
