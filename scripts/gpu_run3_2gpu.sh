set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_search_gpu.py -x -q -m gpu -k "fetch or sharded_entry or gemm_path_parity" > gpurun_out/t_quick.log 2>&1; echo "quick rc=$?"; tail -3 gpurun_out/t_quick.log
timeout 600 $TR tests/sharded_check_torchrun.py > gpurun_out/sharded_check_2.log 2>&1; echo "sharded_check rc=$?"; tail -4 gpurun_out/sharded_check_2.log
timeout 900 $TR tests/sharded_store_check_torchrun.py > gpurun_out/sharded_store_check_2.log 2>&1; echo "store_check rc=$?"; tail -6 gpurun_out/sharded_store_check_2.log
timeout 900 $TR bench.py --gpus 2 > gpurun_out/b_2gpu.log 2>&1; echo "bench2 rc=$?"; tail -1 gpurun_out/b_2gpu.log | cut -c1-3000
timeout 600 $TR benchmarks/sharded_store_bench.py 1000000 768 bf16 > gpurun_out/store_bench_2.log 2>&1; echo "store_bench rc=$?"; tail -8 gpurun_out/store_bench_2.log
