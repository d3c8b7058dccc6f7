set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_fullsize_gpu.py > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/t_all.log | cut -c1-300
timeout 600 python bench.py --rows 1250000 --no-cpu-baseline --no-batched --steps 200 > gpurun_out/b_small.log 2>&1; echo "bsmall rc=$?"; tail -1 gpurun_out/b_small.log | cut -c1-1800
timeout 600 python bench.py --no-cpu-baseline --no-batched > gpurun_out/b_full.log 2>&1; echo "bfull rc=$?"; tail -1 gpurun_out/b_full.log | cut -c1-2200
