set -x
mkdir -p gpurun_out
timeout 600 python benchmarks/configs.py c1 > gpurun_out/configs_c1.log 2>&1; echo "c1 rc=$?"; tail -1 gpurun_out/configs_c1.log | cut -c1-300
