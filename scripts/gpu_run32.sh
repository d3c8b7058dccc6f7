set -x
mkdir -p gpurun_out
timeout 300 python benchmarks/encoder_probe.py 64 512 3 > gpurun_out/enc_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/enc_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 48 -c 4 -o gpurun_out/r02_linear_pair python benchmarks/encoder_probe.py 64 512 2 > gpurun_out/enc_ncu2.log 2>&1; echo "ncu2 rc=$?"
