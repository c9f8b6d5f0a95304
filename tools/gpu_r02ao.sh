#!/bin/bash
# walk: pool slices cut from per-warp chunks (no atomics round trip per node)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02ao_pytest_gpu.log
probe() { # name, n, kind, env...
  local name=$1; local n=$2; local kind=$3; shift; shift; shift
  env "$@" timeout 600 python tools/step_probe.py --n $n --kind $kind --steps 3 > gpurun_out/r02ao_probe_${n}_${kind}_$name.json 2> gpurun_out/r02ao_probe_${n}_${kind}_$name.err
  tail -2 gpurun_out/r02ao_probe_${n}_${kind}_$name.err
  python -c "
import json; j=json.load(open('gpurun_out/r02ao_probe_${n}_${kind}_$name.json')); r=j['resident']; print('$name $n $kind resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pairs', j['pc_pairs'], j['pp_pairs'])"
}
probe chunk 16777216 uniform X=1
probe chunk 4194304 clustered X=1
CB200_EMULATE_RANK=3/8 timeout 600 python tools/step_probe.py --n 16777216 --steps 3 > gpurun_out/r02ao_probe_rank3of8.json 2>/dev/null
python -c "
import json; j=json.load(open('gpurun_out/r02ao_probe_rank3of8.json')); r=j['resident']; print('rank 3/8', round(r['ms_per_step'],3), r['rank_phases_ms'])"
