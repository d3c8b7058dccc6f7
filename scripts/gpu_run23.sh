set -x
mkdir -p gpurun_out
timeout 900 python bench.py --no-c5 > gpurun_out/bench1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench1.log | cut -c1-200
timeout 900 python bench.py --no-c5 --no-batched --steps 200 > gpurun_out/bench1b.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench1b.log | cut -c1-200
