set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -x -q -m gpu -s > gpurun_out/t_full.log 2>&1; echo "fullsize rc=$?"; grep -v "^$" gpurun_out/t_full.log | tail -25 | cut -c1-400
timeout 900 python bench.py > gpurun_out/b_full.log 2>&1; echo "bfull rc=$?"; tail -1 gpurun_out/b_full.log | cut -c1-4000
timeout 300 python bench.py --impl reference > gpurun_out/b_ref.log 2>&1; echo "bref rc=$?"; tail -1 gpurun_out/b_ref.log | cut -c1-1500
timeout 600 python benchmarks/q_sweep.py 10000000 10 > gpurun_out/q_sweep_k10.jsonl 2>&1; echo "qsweep rc=$?"; tail -14 gpurun_out/q_sweep_k10.jsonl | cut -c1-260
