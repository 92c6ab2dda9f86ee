mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_umma_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo TEST FAILED; exit 1; fi
BENCH_REUSE_CONFIGS=one-tile,one-tile-cl2,one-tile-cl4 timeout 600 python profiles/bench_reuse.py > gpurun_out/r02_bench_reuse_cluster.txt 2>&1; cat gpurun_out/r02_bench_reuse_cluster.txt
for cl in 1 2 4; do MSMC_REUSE_CLUSTER=$cl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('cluster=$cl', d['ms_per_step'], d['kernel_families']['msmc_conv_forward_umma_reuse'])"; done
