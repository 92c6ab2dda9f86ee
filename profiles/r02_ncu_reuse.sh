# ncu --set full of the tap-reuse kernels (round-1 one-tile kernel vs the persistent kernel) on single shapes
mkdir -p gpurun_out
cap() { # tag case env regex
  env $3 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$4" --launch-skip 3 --launch-count 1 -f -o /tmp/r02_$1 python profiles/bench_reuse.py $2 > gpurun_out/ncu_$1.log 2>&1
  tail -2 gpurun_out/ncu_$1.log
  ncu -i /tmp/r02_$1.ncu-rep --page raw --csv > gpurun_out/r02_$1_raw.csv 2>/dev/null
  ncu -i /tmp/r02_$1.ncu-rep --page details > gpurun_out/r02_$1_details.txt 2>/dev/null
  ncu -i /tmp/r02_$1.ncu-rep --page source --csv --print-kernel-base demangled > gpurun_out/r02_$1_source.csv 2>/dev/null
  rm -f /tmp/r02_$1.ncu-rep
}
cap ffn2_onetile ffn2 MSMC_X=1 'conv_umma_reuse_kernel'
cap ffn2_persist ffn2 MSMC_X=1 'conv_reuse_persist_kernel'
cap mrf32_persist mrf32k11 MSMC_X=1 'conv_reuse_persist_kernel'
cap mrf32_onetile mrf32k11 MSMC_X=1 'conv_umma_reuse_kernel'
ls -la gpurun_out | tail -20
