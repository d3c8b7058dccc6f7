set -x
mkdir -p gpurun_out
timeout 900 python benchmarks/encoder_bench.py > gpurun_out/encoder_bench.jsonl 2> gpurun_out/encoder_bench.err; echo "encbench rc=$?"; cat gpurun_out/encoder_bench.jsonl | cut -c1-700; tail -3 gpurun_out/encoder_bench.err
timeout 600 python benchmarks/configs.py c1 c2 > gpurun_out/configs_c1c2.log 2>&1; echo "configs rc=$?"; tail -12 gpurun_out/configs_c1c2.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/b_full.log 2>&1; echo "bfull rc=$?"; tail -1 gpurun_out/b_full.log | cut -c1-200
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batched --no-adapter > gpurun_out/plain_for_ncu.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/r02_scan_fused python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-batched --no-adapter > gpurun_out/ncu_scan.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_scan.log
