set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_search_gpu.py -x -q -m gpu > gpurun_out/t_search.log 2>&1; echo "search rc=$?" | tee -a gpurun_out/t_search.log
tail -5 gpurun_out/t_search.log
timeout 600 python bench.py --rows 1250000 --no-cpu-baseline --no-batched --steps 200 > gpurun_out/b_small.log 2>&1; echo "bsmall rc=$?"
tail -1 gpurun_out/b_small.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/b_full.log 2>&1; echo "bfull rc=$?"
tail -1 gpurun_out/b_full.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_search_gpu.py > gpurun_out/t_rest.log 2>&1; echo "rest rc=$?" | tee -a gpurun_out/t_rest.log
tail -8 gpurun_out/t_rest.log
