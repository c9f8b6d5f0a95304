#!/bin/bash
# A/B bench of library variants: tools/gpu_ab.sh lib1.so lib2.so ...
for l in "$@"; do
  echo "== $l"
  CB200_LIB=$l timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 3 --large-n 0 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'ms':round(j['ms_per_step'],4),'pc_ms':round(j['kernels']['pc_ms'],4),'pp_ms':round(j['kernels']['pp_ms'],4),'frac':round(j['roofline']['frac'],4)}))"
done
