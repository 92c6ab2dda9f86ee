mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_modules_gpu.py -m gpu -q -p no:cacheprovider -k "inference_path or mel_extraction" 2>&1 | tail -15
( time CUDA_VISIBLE_DEVICES=0 timeout 600 python msmc-tts_b200/train.py -c msmc-tts_b200/examples/csmsc/msmc_vq_gan_synthetic.yaml ) 2>&1 | grep -E "ms/step|done|real|Error|error" | tail -8
