# Round-1 final captures of one eager train step (profiles/run_step.py), run on the GPU box through gpurun:
#  1. per-launch duration + DRAM bytes of EVERY kernel of the step  -> gpurun_out/step_traffic.csv
#  2. ncu --set full of the heaviest kernel templates, exported to CSV / text on the box (.ncu-rep is 55+ MB)
mkdir -p gpurun_out
(time timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none --csv --log-file gpurun_out/step_traffic.csv python profiles/run_step.py) > gpurun_out/ncu_traffic.log 2>&1
tail -4 gpurun_out/ncu_traffic.log
cap() { # name regex skip count
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o /tmp/r01f_$1 python profiles/run_step.py > gpurun_out/ncuf_$1.log 2>&1
  tail -1 gpurun_out/ncuf_$1.log
  ncu -i /tmp/r01f_$1.ncu-rep --page raw --csv > gpurun_out/r01f_$1_raw.csv 2>/dev/null
  ncu -i /tmp/r01f_$1.ncu-rep --page details > gpurun_out/r01f_$1_details.txt 2>/dev/null
  rm -f /tmp/r01f_$1.ncu-rep
}
cap reuse128 'conv_umma_reuse_kernel<.int.128, .bool.1, .int.2, .int.3, .int.0>' 6 3
cap wgrad128 'conv_wgrad_umma_kernel<.int.128, .bool.1, .int.3, .int.0' 4 3
cap vqcluster 'vq_search_cluster_kernel' 0 2
du -sh gpurun_out
