#!/bin/bash
# thin Ewald launch co-resident with the walk (ewald_pull_kernel): A/B over side CTAs and walk CTAs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02p_pytest_gpu.log
probe() { # name, n, kind, env...
  local name=$1; local n=$2; local kind=$3; shift; shift; shift
  env "$@" timeout 600 python tools/step_probe.py --n $n --kind $kind --steps 3 > gpurun_out/r02p_probe_${n}_${kind}_$name.json 2> gpurun_out/r02p_probe_${n}_${kind}_$name.err
  tail -2 gpurun_out/r02p_probe_${n}_${kind}_$name.err
  python -c "
import json; j=json.load(open('gpurun_out/r02p_probe_${n}_${kind}_$name.json')); r=j['resident']; print('$name $n $kind resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'ew', round(r['rank_ewald_ms'],2))"
}
probe side0 16777216 uniform CB200_EWALD_SIDE_CTAS=0
probe side1w3 16777216 uniform CB200_EWALD_SIDE_CTAS=1 CB200_WALK_CTAS=3
probe side1w4 16777216 uniform CB200_EWALD_SIDE_CTAS=1 CB200_WALK_CTAS=4
probe side2w2 16777216 uniform CB200_EWALD_SIDE_CTAS=2 CB200_WALK_CTAS=2
probe side2w3 16777216 uniform CB200_EWALD_SIDE_CTAS=2 CB200_WALK_CTAS=3
probe side1w3 4194304 uniform CB200_EWALD_SIDE_CTAS=1 CB200_WALK_CTAS=3
probe side0 4194304 uniform CB200_EWALD_SIDE_CTAS=0
probe side1w3 4194304 clustered CB200_EWALD_SIDE_CTAS=1 CB200_WALK_CTAS=3
