#!/bin/bash
# ncu --set full of the leaf walk level in the final state (4.2 M: 22 level launches + 1 ranges launch per step)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"walk_level_kernel" -s 129 -c 1 -f -o gpurun_out/r02an_prof_walk_leaf_4M \
  python tools/step_probe.py --n 4194304 --steps 1 > gpurun_out/r02an_prof_walk.log 2>&1
tail -2 gpurun_out/r02an_prof_walk.log | cut -c1-300
