set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_vq_umma_gpu.py 2>&1 | tail -25 > gpurun_out/r02_pytest1.log
cat gpurun_out/r02_pytest1.log
timeout 300 python -m pytest tests/test_vq_umma_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r02_pytest_vqumma.log
cat gpurun_out/r02_pytest_vqumma.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 > gpurun_out/r02_smoke.log
cat gpurun_out/r02_smoke.log
MSMC_BENCH_DUMP=r02_shapes_a.txt timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.log 2>&1
tail -c 6000 gpurun_out/r02_bench_a.log
