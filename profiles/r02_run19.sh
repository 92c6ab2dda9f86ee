mkdir -p gpurun_out
timeout 300 python profiles/bench_vq.py > gpurun_out/r02_bench_vq_a.txt 2>&1; cat gpurun_out/r02_bench_vq_a.txt
export BENCH_REUSE_CONFIGS=one-tile
for two in 0 1; do echo "== MSMC_REUSE128_TWO=$two"; MSMC_REUSE128_TWO=$two timeout 300 python profiles/bench_reuse.py mrf128k3 mrf128k11 mrf256k11 ffn1 ffn2 ffn2_60 2>&1 | cut -c1-90; done
for two in 0 1; do MSMC_REUSE128_TWO=$two timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1; done
