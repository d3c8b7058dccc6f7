set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu -s > gpurun_out/t_enc.log 2>&1; echo "enc rc=$?"; grep -v "^$" gpurun_out/t_enc.log | tail -30 | cut -c1-300
