#!/bin/bash
# Round-2 multi-GPU validation, in the order of increasing risk; every step under its own timeout so that a hang costs at most
# that timeout (the 8-GPU hang of round 1 cost 123 GPU-minutes).  Run with:   gpurun --gpus 2 --timeout 420 -- 'bash profiles/multi_gpu_r02.sh 2'
# and only after it is green:                                                  gpurun --gpus 8 --timeout 420 -- 'bash profiles/multi_gpu_r02.sh 8'
set +e
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
t0=$(date +%s)
# 1. parity of the sharded path against a replicated oracle (synchronous routing; two-stream apply is the new default)
timeout 120 $RUN tests/run_sharded_gpu.py > gpurun_out/mg_parity_sync_n$N.log 2>&1
echo "parity sync rc=$? t=$(( $(date +%s) - t0 ))s"; tail -2 gpurun_out/mg_parity_sync_n$N.log
# 2. weak-scaling bench, synchronous routing (the published configuration)
timeout 150 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_sync.json 2> gpurun_out/bench_n${N}_sync.err
echo "bench sync rc=$? t=$(( $(date +%s) - t0 ))s"; head -c 300 gpurun_out/bench_n${N}_sync.json; echo
# 3. the order-safe pipelined router: parity first (small, 60 s collective timeout inside), then the bench
DSA_DIST_PIPELINE=1 timeout 90 $RUN tests/run_sharded_gpu.py > gpurun_out/mg_parity_pipe_n$N.log 2>&1
rc=$?
echo "parity pipelined rc=$rc t=$(( $(date +%s) - t0 ))s"; tail -2 gpurun_out/mg_parity_pipe_n$N.log
if [ $rc -eq 0 ]; then
    DSA_DIST_PIPELINE=1 timeout 150 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_pipe.json 2> gpurun_out/bench_n${N}_pipe.err
    echo "bench pipelined rc=$? t=$(( $(date +%s) - t0 ))s"; head -c 300 gpurun_out/bench_n${N}_pipe.json; echo
else
    echo "pipelined parity failed or hung: bench skipped"
fi
