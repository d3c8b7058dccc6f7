set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
timeout 400 $TR bench.py --gpus 4 > gpurun_out/b_4gpu.log 2>&1; echo "bench4 rc=$?"; tail -1 gpurun_out/b_4gpu.log | cut -c1-200
