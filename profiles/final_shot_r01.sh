#!/bin/bash
# Round-1 last GPU call (4 GPU-minutes left): validates the default path, then runs the opt-in experimental switches
# (bit-identity against the default, parity suite, timing).  Everything is wrapped in its own timeout; logs -> gpurun_out/.
set +e
mkdir -p gpurun_out
SW="DSA_SPMV_BULK=1 DSA_TWO_STREAMS=1 DSA_SCAN_ONEPASS=1"
t0=$(date +%s)
timeout 100 python -m pytest tests -q -m gpu > gpurun_out/shot_default_tests.log 2>&1
echo "default tests rc=$? t=$(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/shot_default_tests.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/shot_smoke.log 2>&1
echo "smoke rc=$? t=$(( $(date +%s) - t0 ))s"; tail -2 gpurun_out/shot_smoke.log
env $SW timeout 100 python -m pytest tests -q -m gpu > gpurun_out/shot_allsw_tests.log 2>&1
echo "all-switch tests rc=$? t=$(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/shot_allsw_tests.log
DSA_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_zz_experimental.py -q -m gpu -s -k "update_switches" > gpurun_out/shot_exp_update.log 2>&1
echo "exp update rc=$? t=$(( $(date +%s) - t0 ))s"; grep -h "variant\|passed\|failed\|Error" gpurun_out/shot_exp_update.log | tail -12
DSA_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_zz_experimental.py -q -m gpu -s -k "spmv_bulk and (size0 or size2)" > gpurun_out/shot_exp_spmv.log 2>&1
echo "exp spmv rc=$? t=$(( $(date +%s) - t0 ))s"; grep -h "variant\|passed\|failed\|Error" gpurun_out/shot_exp_spmv.log | tail -16
env $SW timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_allsw.json 2> gpurun_out/bench_allsw.err
echo "bench all-switch rc=$? t=$(( $(date +%s) - t0 ))s"; head -c 600 gpurun_out/bench_allsw.json
