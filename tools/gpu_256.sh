#!/bin/bash
# 256^3 on one GPU (SURVEY config C3 and the C4 recipe), tree and lists on the device
mkdir -p gpurun_out
for kind in uniform clustered; do
  timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 --large-n 16777216 --large-kind $kind > gpurun_out/bench_256_$kind.json 2> gpurun_out/bench_256_$kind.err
  python -c "
import json;j=json.load(open('gpurun_out/bench_256_$kind.json'))['large_box'];json.dump(j,open('gpurun_out/large_256_${kind}_r01i.json','w'));print(json.dumps(j))"
done
