set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/sharded_check_torchrun.py > gpurun_out/sharded_check_8.log 2>&1; echo "sharded_check rc=$?"; tail -4 gpurun_out/sharded_check_8.log
timeout 900 $TR bench.py --gpus 8 > gpurun_out/b_8gpu.log 2>&1; echo "bench8 rc=$?"; tail -1 gpurun_out/b_8gpu.log | cut -c1-6000
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 900 $TR4 bench.py --gpus 4 --no-c5 > gpurun_out/b_4gpu.log 2>&1; echo "bench4 rc=$?"; tail -1 gpurun_out/b_4gpu.log | cut -c1-3000
