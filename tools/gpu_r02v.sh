#!/bin/bash
# ncu --set full of the big moment-build levels at 4 M (launch 110 = deepest level of the sixth step)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"build_moments" -s 111 -c 4 -f -o gpurun_out/r02v_prof_moments_4M \
  python tools/step_probe.py --n 4194304 --steps 1 > gpurun_out/r02v_prof_moments.log 2>&1
tail -2 gpurun_out/r02v_prof_moments.log; ls -la gpurun_out
