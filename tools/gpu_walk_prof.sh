#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:walk_level -s 19 -c 1 -f -o gpurun_out/prof_walk_leaf \
  python tools/bench_device_tree.py --n 4194304 --steps 1 > /dev/null 2>&1
ls -la gpurun_out/prof_walk_leaf.ncu-rep
