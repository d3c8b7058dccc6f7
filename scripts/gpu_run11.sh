set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu -s > gpurun_out/t_enc.log 2>&1; echo "enc rc=$?"; grep -v "^$" gpurun_out/t_enc.log | tail -14 | cut -c1-300
timeout 900 python benchmarks/encoder_bench.py > gpurun_out/encoder_bench.jsonl 2> gpurun_out/encoder_bench.err; echo "encbench rc=$?"; cat gpurun_out/encoder_bench.jsonl | cut -c1-360
timeout 300 python benchmarks/encoder_probe.py 64 512 3 > gpurun_out/enc_probe_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"linear_kernel|attention_kernel|add_ln|embed_ln|pool_kernel" --csv --log-file gpurun_out/r02_launches_encoder_b64_l512.csv python benchmarks/encoder_probe.py 64 512 2 > gpurun_out/enc_ncu1.log 2>&1; echo "ncu1 rc=$?"
