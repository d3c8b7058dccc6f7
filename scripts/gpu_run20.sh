set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_adapter_gpu.py -x -q -m gpu > gpurun_out/t_search.log 2>&1; echo "search rc=$?"; tail -3 gpurun_out/t_search.log
timeout 900 python bench.py > gpurun_out/bench1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench1.log | cut -c1-3000
