#!/bin/bash
# A/B of the FP64 p-c register tiles: CB200_PC64_VARIANT = 0 <8,2> | 1 <6,3> | 2 <12,2> | 3 <4,3> | 4 5 6 = <4,2> <6,2> <8,2> with two targets evaluated statement by statement (pc_pairN<2>)
# parity of a variant: CB200_PC64_VARIANT=4 python -m pytest tests -m gpu -q -k "double or collapse"
for v in "$@"; do
  echo "== variant $v"
  CB200_PC64_VARIANT=$v timeout 300 python tools/resident_probe.py --workload cube300 --double --steps 30 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({k: (round(j[k],4) if isinstance(j[k],float) else j[k]) for k in ('resident_step_ms','pc_ms','pp_ms','ewald_ms','pc_frac_of_fma_peak')}))"
done
