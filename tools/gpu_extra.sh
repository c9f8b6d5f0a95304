#!/bin/bash
# ncu summaries of the device-path kernels (one launch each) + launch list of the raw-particle step
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_rawstep.csv \
  python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --large-n 4194304 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk_level -s 19 -c 1 -f -o gpurun_out/prof_walk_leaf \
  python tools/bench_device_tree.py --n 4194304 --steps 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:emit_fill -s 0 -c 1 -f -o gpurun_out/prof_emit_fill \
  python tools/bench_device_tree.py --n 4194304 --steps 1 > /dev/null 2>&1
ls -la gpurun_out | tail -4
