set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/sharded_store_check_torchrun.py > gpurun_out/sstore2.log 2>&1; echo "sharded store check rc=$?"; tail -4 gpurun_out/sstore2.log
timeout 600 $TR tests/sharded_check_torchrun.py > gpurun_out/scheck2.log 2>&1; echo "sharded check rc=$?"; tail -3 gpurun_out/scheck2.log
timeout 900 $TR bench.py --gpus 2 --no-c5 > gpurun_out/bench2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench2.log | cut -c1-300
