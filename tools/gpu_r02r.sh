#!/bin/bash
# emit_count per thread + flagged kernel; Ewald forked behind the walk's level kernels (CB200_EWALD_AT=1) vs in front
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02r_pytest_gpu.log
probe() { # name, n, kind, env...
  local name=$1; local n=$2; local kind=$3; shift; shift; shift
  env "$@" timeout 600 python tools/step_probe.py --n $n --kind $kind --steps 3 > gpurun_out/r02r_probe_${n}_${kind}_$name.json 2> gpurun_out/r02r_probe_${n}_${kind}_$name.err
  tail -2 gpurun_out/r02r_probe_${n}_${kind}_$name.err
  python -c "
import json; j=json.load(open('gpurun_out/r02r_probe_${n}_${kind}_$name.json')); r=j['resident']; print('$name $n $kind resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pairs', j['pc_pairs'], j['pp_pairs'])"
}
probe at0 16777216 uniform CB200_EWALD_AT=0
probe at1 16777216 uniform CB200_EWALD_AT=1
probe at0 4194304 clustered CB200_EWALD_AT=0
probe at1 4194304 clustered CB200_EWALD_AT=1
probe at0 4194304 uniform CB200_EWALD_AT=0
probe at1 4194304 uniform CB200_EWALD_AT=1
