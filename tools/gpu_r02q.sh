#!/bin/bash
# emit_fill with the piece table (four loads in flight)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02q_pytest_gpu.log
probe() { # name, n, kind, env...
  local name=$1; local n=$2; local kind=$3; shift; shift; shift
  env "$@" timeout 600 python tools/step_probe.py --n $n --kind $kind --steps 3 > gpurun_out/r02q_probe_${n}_${kind}_$name.json 2> gpurun_out/r02q_probe_${n}_${kind}_$name.err
  tail -2 gpurun_out/r02q_probe_${n}_${kind}_$name.err
  python -c "
import json; j=json.load(open('gpurun_out/r02q_probe_${n}_${kind}_$name.json')); r=j['resident']; print('$name $n $kind resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pairs', j['pc_pairs'], j['pp_pairs'])"
}
probe new 16777216 uniform X=1
probe new 4194304 clustered X=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02q_launches_step_256.csv \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02q_launches.log 2>&1
grep -E "emit_fill|walk_paths|emit_count" gpurun_out/r02q_launches_step_256.csv | tail -3 | cut -d, -f5,10,15
