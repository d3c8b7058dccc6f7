set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu > gpurun_out/t_enc.log 2>&1; echo "enc rc=$?"; tail -3 gpurun_out/t_enc.log
timeout 600 python benchmarks/encoder_bench.py > gpurun_out/enc_bench.log 2>&1; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"linear_kernel|attention_kernel|add_ln|embed_ln|pool_kernel" --csv --log-file gpurun_out/r02_launches_encoder_b64_l512_v5.csv python benchmarks/encoder_probe.py 64 512 2 > gpurun_out/enc_ncu1.log 2>&1; echo "ncu1 rc=$?"
