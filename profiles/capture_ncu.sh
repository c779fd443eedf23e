#!/bin/bash
# One-command ncu capture of the config-2 step (run ON THE GPU BOX through gpurun, ~40 s):
#   gpurun --timeout 120 -- 'bash profiles/capture_ncu.sh r02a'
# then, back in the container:
#   python profiles/summarize_launches.py gpurun_out/launches_r02a.csv --steps 2 > profiles/launches_r02a.md
#   python profiles/summarize_full.py gpurun_out/full_r02a.ncu-rep r02a      # -> profiles/ncu_full_r02a.md + ncu_traffic.json
# Both passes profile only the timed steps of profiles/run_c2_steps.py (cudaProfilerStart/Stop around them; the build and
# the warm-up are skipped), i.e. exactly the kernels of bench.py's step, without the torch import.
set -e
TAG=${1:-rXX}
mkdir -p gpurun_out
# 1. launch list: one cheap metric, every launch of 3 steps
DSA_PROFILE_STEPS=1 timeout 60 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv python profiles/run_c2_steps.py gpurun_out/o.npz 100000 10000000 1000000 3
# 2. full set, with source correlation, ONE step (35 launches x ~40 replays)
DSA_PROFILE_STEPS=1 timeout 200 ncu --profile-from-start off --set full --import-source on --clock-control none \
    -o gpurun_out/full_${TAG} -f python profiles/run_c2_steps.py gpurun_out/o.npz 100000 10000000 1000000 1
ls -la gpurun_out/launches_${TAG}.csv gpurun_out/full_${TAG}.ncu-rep
