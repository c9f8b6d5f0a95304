#!/bin/bash
# ncu --set full of the final kernels at 256^3: p-c, p-p, the leaf walk level, emit_fill (one launch each)
mkdir -p gpurun_out
# 6 steps per probe run (3 warm, e2e, upload, resident): the last launch of each kernel family
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cell_list_x2" -s 5 -c 1 -f -o gpurun_out/r02aa_prof_pc_256 \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02aa_prof_pc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"part_list_stream" -s 5 -c 1 -f -o gpurun_out/r02aa_prof_pp_256 \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02aa_prof_pp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"emit_fill" -s 5 -c 1 -f -o gpurun_out/r02aa_prof_emit_256 \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02aa_prof_emit.log 2>&1
# 25 levels per step: level 21 (the bucket level) of the sixth step = launch 5 * 25 + 21
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"walk_level" -s 146 -c 1 -f -o gpurun_out/r02aa_prof_walk_256 \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02aa_prof_walk.log 2>&1
ls -la gpurun_out
