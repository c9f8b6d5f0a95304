#!/bin/bash
# walk at 5 CTAs/SM (96 registers, 44 KB of shared memory per CTA) against 4 and 6
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02ak_pytest_gpu.log
probe() { # name, n, kind, env...
  local name=$1; local n=$2; local kind=$3; shift; shift; shift
  env "$@" timeout 600 python tools/step_probe.py --n $n --kind $kind --steps 3 > gpurun_out/r02ak_probe_${n}_${kind}_$name.json 2> gpurun_out/r02ak_probe_${n}_${kind}_$name.err
  tail -2 gpurun_out/r02ak_probe_${n}_${kind}_$name.err
  python -c "
import json; j=json.load(open('gpurun_out/r02ak_probe_${n}_${kind}_$name.json')); r=j['resident']; print('$name $n $kind resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pairs', j['pc_pairs'], j['pp_pairs'])"
}
probe w5 16777216 uniform X=1
probe w4 16777216 uniform CB200_LIB=changa_b200/variants/walk4.so
probe w6 16777216 uniform CB200_LIB=changa_b200/variants/walk6.so
probe w5 4194304 clustered X=1
probe w4 4194304 clustered CB200_LIB=changa_b200/variants/walk4.so
