set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python -m pytest tests/test_search_gpu.py -x -q -m gpu > gpurun_out/t_search.log 2>&1; echo "search rc=$?" | tee -a gpurun_out/t_search.log
tail -15 gpurun_out/t_search.log
timeout 600 python bench.py --rows 1250000 --no-cpu-baseline --no-batched --steps 200 > gpurun_out/b_small.log 2>&1; echo "bsmall rc=$?"
tail -2 gpurun_out/b_small.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b_full.log 2>&1; echo "bfull rc=$?"
tail -2 gpurun_out/b_full.log
