mkdir -p gpurun_out
nvidia-smi -L
echo "=== train_dist.py, 2 GPUs, 50 steps"
( time timeout 900 python msmc-tts_b200/train_dist.py -n 2 -s gpurun_out/r02_train_dist_logs -c msmc-tts_b200/examples/csmsc/msmc_vq_gan_synthetic.yaml ) > gpurun_out/r02_train_dist_2gpu.log 2>&1
echo "exit code $?" >> gpurun_out/r02_train_dist_2gpu.log
tail -15 gpurun_out/r02_train_dist_2gpu.log
echo "=== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.log 2>&1
echo "exit code $?"
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n2.log | head -2; grep -o '"value": [0-9.]*' gpurun_out/r02_bench_n2.log | head -1
tail -3 gpurun_out/r02_bench_n2.log | cut -c1-300
echo "=== train.py, 1 GPU, 50 steps"
( time CUDA_VISIBLE_DEVICES=0 timeout 900 python msmc-tts_b200/train.py -c msmc-tts_b200/examples/csmsc/msmc_vq_gan_synthetic.yaml ) > gpurun_out/r02_train_1gpu.log 2>&1
echo "exit code $?" >> gpurun_out/r02_train_1gpu.log
tail -8 gpurun_out/r02_train_1gpu.log
echo "=== bench N=1"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_e.log 2>&1
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_e.log | head -2
