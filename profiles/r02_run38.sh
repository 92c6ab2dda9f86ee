mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r02_final_bench.log 2>&1; tail -c 300 gpurun_out/r02_final_bench.log
bash profiles/r02_final_profile.sh 2>&1 | tail -12
