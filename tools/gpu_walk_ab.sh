#!/bin/bash
# device-walk A/B: list parity tests, then the 4 M boxes with each library given (default lib = "")
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q -k "device or multistep or clustered or raw" 2>&1 | tail -5
for l in "$@"; do
  for kind in uniform clustered; do
    echo "== lib=$l kind=$kind"
    CB200_LIB=$l timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 --large-kind $kind | python -c "import json,sys; j=json.loads(sys.stdin.read())['large_box']; print(json.dumps({'ms':round(j['ms_per_step'],3),'phases':j['rank0_phases_ms']}))"
  done
done
