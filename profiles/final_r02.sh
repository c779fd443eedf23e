#!/bin/bash
# Round-2 closing measurements on ONE B200 (run through gpurun; ~10 minutes, every step under its own timeout):
#   gpurun --timeout 1500 -- 'bash profiles/final_r02.sh'
set +e
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
(timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/final_pytest.log; tail -2 gpurun_out/final_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_bench_reference.json 2> /dev/null; echo "reference rc=$?"
# sanitizers on the final kernels: parity subset + the sharded worker on one rank (routing, push, unpack, device-side counts)
(timeout 500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
    -k "vec_batches or matrix_batches or delete or spmv or build_layout or single_writes" 2>&1 | tail -12) > gpurun_out/final_memcheck.log
# the tile-streamed pipeline (forced on): every test of tests/test_gpu_tile.py under memcheck, the batch tests under racecheck + synccheck
(timeout 500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_tile.py -q -m gpu -x 2>&1 | tail -8) > gpurun_out/final_memcheck_tile.log
echo "memcheck tile: $(tail -1 gpurun_out/final_memcheck_tile.log)"
(timeout 700 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_tile.py -q -m gpu -x 2>&1 | tail -8) > gpurun_out/final_racecheck_tile.log
echo "racecheck tile: $(tail -1 gpurun_out/final_racecheck_tile.log)"
(timeout 400 compute-sanitizer --tool synccheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_tile.py -q -m gpu -x -k "oracle_and_vs_random or hot_leaves" 2>&1 | tail -8) > gpurun_out/final_synccheck_tile.log
echo "synccheck tile: $(tail -1 gpurun_out/final_synccheck_tile.log)"
echo "memcheck: $(tail -1 gpurun_out/final_memcheck.log)"
(RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29571 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 \
    --print-limit 20 python tests/run_sharded_gpu.py 2>&1 | tail -8) > gpurun_out/final_memcheck_sharded.log
echo "memcheck sharded: $(tail -1 gpurun_out/final_memcheck_sharded.log)"
(timeout 700 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
    -k "matrix_batches or delete_columns or spmv_all or single_writes" 2>&1 | tail -8) > gpurun_out/final_racecheck.log
echo "racecheck: $(tail -1 gpurun_out/final_racecheck.log)"
bash profiles/capture_ncu.sh r02e > gpurun_out/final_capture.log 2>&1; tail -2 gpurun_out/final_capture.log
