#!/bin/bash
# launch list of one 256^3 step at the final state of the round (gpu__time_duration, serialised)
mkdir -p gpurun_out
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02at_launches_step_256.csv \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02at_launches.log 2>&1
tail -2 gpurun_out/r02at_launches.log | cut -c1-300
python tools/launch_summary.py gpurun_out/r02at_launches_step_256.csv gpurun_out/r02at_launch_summary_256.json
