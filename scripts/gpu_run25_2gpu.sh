set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR tests/sharded_store_check_torchrun.py > gpurun_out/sstore2.log 2>&1; echo "sharded store check rc=$?"; grep -c "OK" gpurun_out/sstore2.log; tail -3 gpurun_out/sstore2.log | cut -c1-300
