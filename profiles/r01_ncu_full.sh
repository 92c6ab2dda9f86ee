# ncu --set full captures of the heaviest conv kernel variants inside one eager train step (profiles/run_step.py).
# The .ncu-rep files (55+ MB each with --import-source) are exported to CSV on the box and deleted.
mkdir -p gpurun_out
cap() { # name regex skip count want_source
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o /tmp/r01_$1 python profiles/run_step.py > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
  ncu -i /tmp/r01_$1.ncu-rep --page raw --csv > gpurun_out/r01_$1_raw.csv 2>/dev/null
  ncu -i /tmp/r01_$1.ncu-rep --page details > gpurun_out/r01_$1_details.txt 2>/dev/null
  if [ "$5" = "1" ]; then ncu -i /tmp/r01_$1.ncu-rep --page source --csv --print-kernel-base demangled > gpurun_out/r01_$1_source.csv 2>/dev/null; fi
  rm -f /tmp/r01_$1.ncu-rep
}
cap reuse32 'conv_umma_reuse_kernel<.int.32, .bool.1, .int.1, .int.6, .int.0>' 4 3 1
cap wgrad128 'conv_wgrad_umma_kernel<.int.128, .bool.1, .int.3, .int.0' 4 3 1
cap wgradreuse64 'conv_wgrad_reuse_kernel<.int.64' 4 3 0
cap umma32 'conv_umma_kernel<.int.32, .bool.1, .int.2, .int.0' 10 3 0
du -sh gpurun_out; ls -la gpurun_out
