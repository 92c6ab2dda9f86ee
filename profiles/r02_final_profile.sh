# Round-2 final captures of one eager train step (profiles/run_step.py), run on the GPU box through gpurun:
#  1. per-launch duration + DRAM bytes of EVERY kernel of the step  -> gpurun_out/r02_step_traffic.csv
#  2. ncu --set full of the heaviest kernel templates, exported to CSV / text on the box
#  3. CUPTI timeline of the graph-replayed step (profiles/trace_step.py)
mkdir -p gpurun_out
(time timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none --csv --log-file gpurun_out/r02_step_traffic.csv python profiles/run_step.py) > gpurun_out/r02_ncu_traffic.log 2>&1
tail -3 gpurun_out/r02_ncu_traffic.log
python profiles/summarize_launches.py gpurun_out/r02_step_traffic.csv 60 > gpurun_out/r02_final_step_launches_summary.txt 2>&1
python profiles/family_traffic.py gpurun_out/r02_step_traffic.csv gpurun_out/r02_family_traffic.json > gpurun_out/r02_final_step_families.txt 2>&1
head -24 gpurun_out/r02_final_step_families.txt
cap() { # name regex skip count
  timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o /tmp/r02f_$1 python profiles/run_step.py > gpurun_out/r02_ncuf_$1.log 2>&1
  tail -1 gpurun_out/r02_ncuf_$1.log
  ncu -i /tmp/r02f_$1.ncu-rep --page raw --csv > gpurun_out/r02f_$1_raw.csv 2>/dev/null
  python profiles/ncu_raw_summary.py gpurun_out/r02f_$1_raw.csv tc_wavefronts > gpurun_out/r02_final_ncu_full_$1_summary.txt 2>/dev/null
  rm -f /tmp/r02f_$1.ncu-rep
}
cap wgrad128 'conv_wgrad_umma_kernel<.int.128, .bool.1, .int.3, .int.0' 4 3
cap wgradreuse128 'conv_wgrad_reuse_kernel<.int.128' 2 3
cap reuse128 'conv_umma_reuse_kernel<.int.128, .bool.1, .int.2, .int.3, .int.0>' 6 3
cap persist64 'conv_reuse_persist_kernel<.int.64' 2 2
# the two-phase tensor-core VQ search alone at 2^18 rows, K = 256 and K = 64
for K in 256 64; do
  timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:vq_search_umma" --launch-skip 2 --launch-count 1 -f -o /tmp/r02f_vq$K python profiles/run_vq_case.py 262144 $K umma > gpurun_out/r02_ncuf_vq$K.log 2>&1
  tail -1 gpurun_out/r02_ncuf_vq$K.log
  ncu -i /tmp/r02f_vq$K.ncu-rep --page raw --csv > gpurun_out/r02f_vq${K}_raw.csv 2>/dev/null
  python profiles/ncu_raw_summary.py gpurun_out/r02f_vq${K}_raw.csv > gpurun_out/r02_final_ncu_full_vqumma${K}_summary.txt 2>/dev/null
  rm -f /tmp/r02f_vq$K.ncu-rep
done
gzip -f gpurun_out/r02_step_traffic.csv
timeout 600 python profiles/trace_step.py r02_final 2>&1 | tail -1
gzip -f gpurun_out/r02_final_timeline.csv
du -sh gpurun_out
