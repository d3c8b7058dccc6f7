set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_search_gpu.py -x -q -m gpu -k "pipelined_sharded or sharded_entry" > gpurun_out/t_new.log 2>&1; echo "new rc=$?"; tail -3 gpurun_out/t_new.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-adapter > gpurun_out/bench_plain_for_ncu.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 -k regex:"scan_topk|gemm_topk|finalize_kernel|prep_|exchange_merge|upsert_kernel" --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-adapter > gpurun_out/ncu_l.log 2>&1; echo "ncu-launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 4 -c 1 -o gpurun_out/r02_gemm_pair python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-adapter > gpurun_out/ncu_g.log 2>&1; echo "ncu-gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:finalize_kernel -s 4 -c 1 -o gpurun_out/r02_finalize python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-adapter > gpurun_out/ncu_f.log 2>&1; echo "ncu-fin rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:exchange_merge -c 1 -o gpurun_out/r02_exchange python -m pytest tests/test_search_gpu.py -x -q -m gpu -k "sharded_entry and 130" > gpurun_out/ncu_x.log 2>&1; echo "ncu-x rc=$?"
