set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
timeout 600 $TR bench.py --gpus 8 > gpurun_out/b_8gpu.log 2>&1; echo "bench8 rc=$?"; tail -1 gpurun_out/b_8gpu.log | cut -c1-300
timeout 200 $TR tests/sharded_store_check_torchrun.py > gpurun_out/sstore8.log 2>&1; echo "sharded store check rc=$?"; grep "OK" gpurun_out/sstore8.log
