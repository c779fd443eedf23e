#!/bin/bash
# Where the tile-streamed pipeline overtakes the random-access one at config 2's shape (capacity 2^24 per orientation):
# the same device-resident steps with DSA_TILE=0 (never) and DSA_TILE=2 (always), for several batch sizes.
# The two runs of a size must print the same layout digests.
#   gpurun --timeout 600 -- 'bash profiles/tile_crossover.sh > gpurun_out/tile_crossover.log'
for b in 65536 131072 262144 524288 1000000 2000000; do
  for mode in 0 2; do
    DSA_TILE=$mode timeout 120 python profiles/run_c2_steps.py gpurun_out/x_${b}_${mode}.npz 100000 10000000 $b 10 2>&1 | tail -1
  done
done
