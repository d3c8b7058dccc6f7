set -x
mkdir -p gpurun_out
timeout 300 python benchmarks/encoder_probe.py 64 512 3 > gpurun_out/enc_probe_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"linear_kernel|attention_kernel|add_ln|embed_ln|pool_kernel" --csv --log-file gpurun_out/r02_launches_encoder_b64_l512.csv python benchmarks/encoder_probe.py 64 512 2 > gpurun_out/enc_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 48 -c 4 -o gpurun_out/r02_linear python benchmarks/encoder_probe.py 64 512 2 > gpurun_out/enc_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 12 -c 1 -o gpurun_out/r02_attention python benchmarks/encoder_probe.py 64 512 2 > gpurun_out/enc_ncu3.log 2>&1; echo "ncu3 rc=$?"
tail -2 gpurun_out/enc_probe_plain.log
