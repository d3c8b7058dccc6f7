set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_all.log | cut -c1-300
timeout 900 python benchmarks/encoder_bench.py > gpurun_out/encoder_bench.jsonl 2> gpurun_out/encoder_bench.err; echo "encbench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/encoder_bench.jsonl'):
    d=json.loads(l); print(d["B"],d["L"],d["ragged"],"ours",round(d["device_ms"],2),"hf fp32",round(d["transformers_fp32_ms"],1),"hf bf16",round(d["transformers_bf16_ms"],2),"cos",round(d["min_cosine_vs_transformers_fp32"],6))
PY
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
