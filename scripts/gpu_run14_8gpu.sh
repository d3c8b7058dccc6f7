set -x
mkdir -p gpurun_out
for N in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
  timeout 900 $TR bench.py --gpus $N > gpurun_out/b_${N}gpu.log 2>&1; echo "bench$N rc=$?"; tail -1 gpurun_out/b_${N}gpu.log | cut -c1-300
done
