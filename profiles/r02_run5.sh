mkdir -p gpurun_out
for bn in 32 64 128; do echo "== MSMC_FORCE_BN=$bn"; MSMC_FORCE_BN=$bn timeout 600 python profiles/bench_reuse.py mrf128k3 mrf128k11 mrf256k11 ffn1 ffn2 ffn2_60 2>&1 | cut -c1-140; done > gpurun_out/r02_bench_reuse_bn.txt
cat gpurun_out/r02_bench_reuse_bn.txt
