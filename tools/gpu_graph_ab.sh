#!/bin/bash
for f in "" "--no-graph"; do
timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 20 --large-n 0 $f | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(json.dumps({'launch':j['config']['launch'],'ms':round(j['ms_per_step'],4),'k':{a:round(b,4) for a,b in j['kernels'].items() if a.endswith('_ms')},'frac':round(j['roofline']['frac'],4),'launches':j['gpu_launches']}))"
done
