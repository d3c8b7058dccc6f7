set -x
mkdir -p gpurun_out
cd benchmarks && timeout 600 python inline_overlap_probe.py 10000000 > ../gpurun_out/inline_probe.log 2>&1; echo "probe rc=$?"; tail -2 ../gpurun_out/inline_probe.log | cut -c1-2500; cd ..
