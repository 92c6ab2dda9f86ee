mkdir -p gpurun_out
cap() { # name regex skip count
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o /tmp/r02f_$1 python profiles/run_step.py > gpurun_out/r02_ncuf_$1.log 2>&1
  tail -1 gpurun_out/r02_ncuf_$1.log
  ncu -i /tmp/r02f_$1.ncu-rep --page raw --csv > gpurun_out/r02f_$1_raw.csv 2>/dev/null
  python profiles/ncu_raw_summary.py gpurun_out/r02f_$1_raw.csv tc_wavefronts > gpurun_out/r02_final_ncu_full_$1_summary.txt 2>/dev/null
  rm -f /tmp/r02f_$1.ncu-rep
}
cap fwd128 'conv_umma_kernel<.int.128, .bool.1, .int.3, .int.0, .bool.0>' 10 4
grep -E "==|duration|tc_cycles_active|inst_executed.sum|warps_active|waves|dram__bytes" gpurun_out/r02_final_ncu_full_fwd128_summary.txt | head -40
