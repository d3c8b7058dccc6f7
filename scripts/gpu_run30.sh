set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_adapter_gpu.py tests/test_search_rank_gpu.py -x -q -m gpu > gpurun_out/t_ad.log 2>&1; echo "adapter rc=$?"; tail -2 gpurun_out/t_ad.log
timeout 600 python benchmarks/configs.py c1 > gpurun_out/configs_c1.log 2>&1; echo "c1 rc=$?"; tail -1 gpurun_out/configs_c1.log | cut -c1-200
