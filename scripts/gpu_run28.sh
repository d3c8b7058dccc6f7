set -x
mkdir -p gpurun_out
timeout 300 python benchmarks/encoder_probe.py 64 512 3 > gpurun_out/enc_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/enc_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 12 -c 1 -o gpurun_out/r02_attention_v4 python benchmarks/encoder_probe.py 64 512 2 > gpurun_out/enc_ncu3.log 2>&1; echo "ncu3 rc=$?"
