#!/bin/bash
# A/B: p-c with two bodies per basic block (CB200_V_DOUBLE), walk at 3 CTAs/SM without the register cap, p-p at 3 CTAs/SM
mkdir -p gpurun_out
probe() { # name, n, kind, env...
  local name=$1; local n=$2; local kind=$3; shift; shift; shift
  env "$@" timeout 600 python tools/step_probe.py --n $n --kind $kind --steps 3 > gpurun_out/r02z_probe_${n}_${kind}_$name.json 2> gpurun_out/r02z_probe_${n}_${kind}_$name.err
  tail -2 gpurun_out/r02z_probe_${n}_${kind}_$name.err
  python -c "
import json; j=json.load(open('gpurun_out/r02z_probe_${n}_${kind}_$name.json')); r=j['resident']; print('$name $n $kind resident', round(r['ms_per_step'],3), r['rank_phases_ms'])"
}
probe base 16777216 uniform X=1
probe vdouble 16777216 uniform CB200_LIB=changa_b200/variants/vdouble.so
probe walk3 16777216 uniform CB200_LIB=changa_b200/variants/walk3.so
probe pp3 16777216 uniform CB200_PP_VARIANT=3
