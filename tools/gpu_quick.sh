#!/bin/bash
# full GPU suite + headline bench (no CPU baseline)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 20 --large-n 0 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'ms':round(j['ms_per_step'],4),'k':{a:round(b,4) for a,b in j['kernels'].items() if a.endswith('_ms')},'frac':round(j['roofline']['frac'],4),'e2e_ms':round(j['e2e']['ms_per_step'],4)}))"
timeout 200 python bench.py --double --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 20 --large-n 0 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'ms':round(j['ms_per_step'],4),'k':{a:round(b,4) for a,b in j['kernels'].items() if a.endswith('_ms')},'frac':round(j['roofline']['frac'],4),'e2e_ms':round(j['e2e']['ms_per_step'],4)}))"
