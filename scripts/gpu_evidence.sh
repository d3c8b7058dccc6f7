#!/usr/bin/env bash
# How the evidence under profiles/ is produced on a B200 box (run through `gpurun -- 'bash scripts/gpu_evidence.sh <what>'`,
# `gpurun --gpus N` for the multi-GPU legs).  Everything lands in gpurun_out/; the summaries worth keeping are copied into
# profiles/ by hand (named per round).  A number printed by a run under ncu is never a bench value: every ncu pass is preceded
# by the same command without ncu, and only runs when that one exited 0.
#
#   suite      full GPU test suite + smoke + bench.py (N=1) + the reference arm            (~8 min)
#   ncu        launch list of bench.py, then --set full captures of the scan, tensor-core, finalize and exchange kernels
#   encoder    encoder tests, encoder bench, per-kernel launch list of one 64 x 512 forward pass
#   multi N    bench.py over N GPUs + the sharded store / sharded searcher checks
set -x
mkdir -p gpurun_out
what=${1:-suite}
BENCH_NCU="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-adapter"
case "$what" in
suite)
    timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "all rc=$?"; tail -3 gpurun_out/t_all.log
    timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
    timeout 900 python bench.py > gpurun_out/bench1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench1.log | cut -c1-300
    timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-400
    ;;
ncu)
    timeout 600 $BENCH_NCU > gpurun_out/bench_plain_for_ncu.log 2>&1 || { echo "plain bench failed"; exit 1; }
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 \
        -k regex:"scan_topk|gemm_topk|finalize_kernel|prep_|exchange_merge|upsert_kernel" --csv \
        --log-file gpurun_out/launches_bench.csv $BENCH_NCU > gpurun_out/ncu_l.log 2>&1; echo "ncu-launches rc=$?"
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/scan_fused \
        $BENCH_NCU --no-batched > gpurun_out/ncu_s.log 2>&1; echo "ncu-scan rc=$?"
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk -s 4 -c 1 -o gpurun_out/gemm_pair \
        $BENCH_NCU > gpurun_out/ncu_g.log 2>&1; echo "ncu-gemm rc=$?"
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:finalize_kernel -s 4 -c 1 -o gpurun_out/finalize \
        $BENCH_NCU > gpurun_out/ncu_f.log 2>&1; echo "ncu-fin rc=$?"
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:exchange_merge -c 1 -o gpurun_out/exchange \
        python -m pytest tests/test_search_gpu.py -x -q -m gpu -k "sharded_entry and 130" > gpurun_out/ncu_x.log 2>&1; echo "ncu-x rc=$?"
    ;;
encoder)
    timeout 600 python -m pytest tests/test_encoder_gpu.py -x -q -m gpu > gpurun_out/t_enc.log 2>&1; echo "enc rc=$?"; tail -3 gpurun_out/t_enc.log
    timeout 600 python benchmarks/encoder_bench.py > gpurun_out/enc_bench.log 2>&1; echo "bench rc=$?"
    timeout 300 python benchmarks/encoder_probe.py 64 512 3 > gpurun_out/enc_plain.log 2>&1 || { echo "plain probe failed"; exit 1; }
    timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
        -k regex:"linear_kernel|attention_kernel|add_ln|embed_ln|pool_kernel" --csv --log-file gpurun_out/launches_encoder_b64_l512.csv \
        python benchmarks/encoder_probe.py 64 512 2 > gpurun_out/enc_ncu.log 2>&1; echo "ncu rc=$?"
    ;;
multi)
    N=${2:-2}
    TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518"
    timeout 600 $TR bench.py --gpus $N > gpurun_out/b_${N}gpu.log 2>&1; echo "bench$N rc=$?"; tail -1 gpurun_out/b_${N}gpu.log | cut -c1-300
    timeout 300 $TR tests/sharded_check_torchrun.py > gpurun_out/sharded$N.log 2>&1; echo "sharded searcher check rc=$?"; grep -c "OK" gpurun_out/sharded$N.log
    timeout 400 $TR tests/sharded_store_check_torchrun.py > gpurun_out/sstore$N.log 2>&1; echo "sharded store check rc=$?"; grep -c "OK" gpurun_out/sstore$N.log
    timeout 300 $TR benchmarks/sharded_store_bench.py > gpurun_out/sstore_bench$N.log 2>&1; echo "sharded store bench rc=$?"
    ;;
*)
    echo "usage: $0 suite | ncu | encoder | multi N"; exit 2;;
esac
