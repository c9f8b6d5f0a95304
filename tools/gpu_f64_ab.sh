#!/bin/bash
# A/B of the FP64 p-c register tiles: CB200_PC64_VARIANT = 0 <8,2> | 1 <6,3> | 2 <12,2> | 3 <4,3> | 4 5 6 = <4,2> <6,2> <8,2> with two targets evaluated statement by statement (pc_pairN<2>)
# parity of a variant: CB200_PC64_VARIANT=4 python -m pytest tests -m gpu -q -k "double or collapse"
for v in "$@"; do
  echo "== variant $v"
  CB200_PC64_VARIANT=$v timeout 300 python bench.py --double --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 3 --large-n 0 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'ms':round(j['ms_per_step'],4),'pc_ms':round(j['kernels']['pc_ms'],4),'pp_ms':round(j['kernels']['pp_ms'],4),'ew_ms':round(j['kernels']['ewald_ms'],4),'frac':round(j['roofline']['frac'],4)}))"
done
