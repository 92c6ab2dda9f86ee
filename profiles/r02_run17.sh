mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python msmc-tts_b200/train.py -c msmc-tts_b200/examples/csmsc/msmc_vq_gan_synthetic.yaml > gpurun_out/r02_train_1gpu.log 2>&1
grep -E "ms/step|done|Error" gpurun_out/r02_train_1gpu.log | cut -c1-60,320-400 | tail
timeout 600 python msmc-tts_b200/train_dist.py -n 2 -s gpurun_out/r02_train_dist_logs -c msmc-tts_b200/examples/csmsc/msmc_vq_gan_synthetic.yaml > gpurun_out/r02_train_dist_2gpu.log 2>&1; echo "exit $?"
grep -E "ms/step|done|Error" gpurun_out/r02_train_dist_2gpu.log | cut -c1-60,330-420 | tail
