#!/bin/bash
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 | python -c "import json,sys; j=json.loads(sys.stdin.read())['large_box']; print(json.dumps({'ms':round(j['ms_per_step'],3),'phases':j['rank0_phases_ms']}))"
