#!/bin/bash
# RawParticleStep at the cube300 size (110 592 particles, uniform generator): the device path end to end
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 5 --large-n 110592 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'e2e_abi_ms':round(j['e2e']['ms_per_step'],4),'large':j['large_box']}))"
