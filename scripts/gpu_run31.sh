set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu > gpurun_out/t_enc.log 2>&1; echo "enc rc=$?"; tail -3 gpurun_out/t_enc.log
timeout 600 python benchmarks/encoder_bench.py > gpurun_out/enc_bench.log 2>&1; echo "bench rc=$?"; grep -o '"B": [0-9]*, "L": [0-9]*, "ragged": [a-z]*, "non_pad_tokens": [0-9]*, "device_ms": [0-9.]*' gpurun_out/enc_bench.log; grep -o '"transformers_bf16_ms": [0-9.]*, "min_cosine[^}]*' gpurun_out/enc_bench.log
