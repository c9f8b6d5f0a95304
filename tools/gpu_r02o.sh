#!/bin/bash
# walk_node_fast with head flushes: tests, probes (uniform / clustered), give-up stats, ncu of the leaf levels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02o_pytest_gpu.log
probe() { # name, n, kind, env...
  local name=$1; local n=$2; local kind=$3; shift; shift; shift
  env "$@" timeout 600 python tools/step_probe.py --n $n --kind $kind --steps 3 > gpurun_out/r02o_probe_${n}_${kind}_$name.json 2> gpurun_out/r02o_probe_${n}_${kind}_$name.err
  tail -2 gpurun_out/r02o_probe_${n}_${kind}_$name.err
  python -c "
import json; j=json.load(open('gpurun_out/r02o_probe_${n}_${kind}_$name.json')); r=j['resident']; print('$name $n $kind resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pairs', j['pc_pairs'], j['pp_pairs'])"
}
probe fast 16777216 uniform X=1
probe fast 4194304 uniform X=1
probe fast 4194304 clustered X=1
probe general 4194304 clustered CB200_WALK_GENERAL=1
probe fast 16777216 clustered X=1
CB200_LIB=changa_b200/variants/stats.so timeout 300 python tools/step_probe.py --n 4194304 --kind clustered --steps 1 2>&1 >/dev/null | sort | uniq -c | tee gpurun_out/r02o_stats_4Mclu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02o_launches_step_256.csv \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02o_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"walk_level" -s 106 -c 3 -f -o gpurun_out/r02o_prof_walk_4M \
  python tools/step_probe.py --n 4194304 --steps 1 > gpurun_out/r02o_prof_walk.log 2>&1
ls -la gpurun_out | tail -3
