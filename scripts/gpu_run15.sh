set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu -s > gpurun_out/t_enc.log 2>&1; echo "enc rc=$?"; grep -v "^$" gpurun_out/t_enc.log | tail -4 | cut -c1-300
timeout 900 python benchmarks/encoder_bench.py > gpurun_out/encoder_bench.jsonl 2> gpurun_out/encoder_bench.err; echo "encbench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/encoder_bench.jsonl'):
    d=json.loads(l); print(d["B"],d["L"],d["ragged"],"ours",round(d["device_ms"],2),"hf fp32",round(d["transformers_fp32_ms"],1),"hf bf16",round(d["transformers_bf16_ms"],2),"cos",round(d["min_cosine_vs_transformers_fp32"],6))
PY
timeout 300 python benchmarks/encoder_probe.py 64 512 3 > gpurun_out/enc_probe_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"linear_kernel|attention_kernel|add_ln|embed_ln|pool_kernel" --csv --log-file gpurun_out/r02_launches_encoder_b64_l512.csv python benchmarks/encoder_probe.py 64 512 2 > gpurun_out/enc_ncu1.log 2>&1; echo "ncu1 rc=$?"
