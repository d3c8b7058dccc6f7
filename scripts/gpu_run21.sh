set -x
mkdir -p gpurun_out
cd benchmarks && timeout 600 python inline_overlap_probe.py 5000000 > ../gpurun_out/inline_probe.log 2>&1; echo "probe rc=$?"; tail -2 ../gpurun_out/inline_probe.log | cut -c1-1500; cd ..
timeout 900 python bench.py --no-c5 > gpurun_out/bench1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench1.log | cut -c1-600
timeout 600 python benchmarks/configs.py c1 > gpurun_out/configs_c1.log 2>&1; echo "c1 rc=$?"; tail -1 gpurun_out/configs_c1.log | cut -c1-400
