#!/bin/bash
# Round-1 very last GPU call (2.3 GPU-minutes left): validates the flipped defaults (two streams + one-pass scan), takes the
# round's final bench line, then measures the DSMEM-gather SpMV variant (mode 8) against flat and the best bulk mode (4).
set +e
mkdir -p gpurun_out
t0=$(date +%s)
timeout 80 python -m pytest tests -q -m gpu > gpurun_out/shot2_tests.log 2>&1
echo "tests rc=$? t=$(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/shot2_tests.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/shot2_smoke.log 2>&1
echo "smoke rc=$? t=$(( $(date +%s) - t0 ))s"; tail -2 gpurun_out/shot2_smoke.log
timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_final2.json 2> gpurun_out/bench_r01_final2.err
echo "bench rc=$? t=$(( $(date +%s) - t0 ))s"; head -c 400 gpurun_out/bench_r01_final2.json; echo
DSA_EXPERIMENTAL=1 DSA_EXP_SPMV_MODES=4,8 timeout 60 python -m pytest tests/test_zz_experimental.py -q -m gpu -s -k "spmv_bulk and (size0 or size2)" > gpurun_out/shot2_exp_spmv.log 2>&1
echo "exp spmv rc=$? t=$(( $(date +%s) - t0 ))s"; grep -h "variant\|passed\|failed\|Error\|error" gpurun_out/shot2_exp_spmv.log | tail -12
