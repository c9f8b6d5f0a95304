#!/bin/bash
# tests with the in-tree library, then the bench with each library given ("" = in-tree)
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for l in "$@"; do
CB200_LIB=$l timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 5 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'lib':'$l','ms':round(j['ms_per_step'],4),'k':{a:round(b,4) for a,b in j['kernels'].items() if a.endswith('_ms')},'large_ms':round(j['large_box']['ms_per_step'],3),'large_pp':round(j['large_box']['rank0_pp_ms'],3),'large_pc':round(j['large_box']['rank0_pc_ms'],3)}))"
done
