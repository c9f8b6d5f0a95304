#!/bin/bash
# ncu --set full of emit_fill (path table), the leaf walk level and the two list kernels at 4 M
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"emit_fill|walk_level" -s 80 -c 30 -f -o gpurun_out/r02l_prof_walk_4M \
  python tools/step_probe.py --n 4194304 --steps 1 > gpurun_out/r02l_prof_walk.log 2>&1
tail -3 gpurun_out/r02l_prof_walk.log
ls -la gpurun_out | tail -3
