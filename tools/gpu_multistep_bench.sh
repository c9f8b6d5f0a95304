#!/bin/bash
for r in 2 4; do
  timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 --large-kind clustered --large-active-rung $r > gpurun_out/bench_multistep_r$r.json 2> gpurun_out/bench_multistep_r$r.err
  python -c "import json;j=json.load(open('gpurun_out/bench_multistep_r$r.json'))['large_box'];print(json.dumps({'rung':$r,'ms':j['ms_per_step'],'multistep':j['multistep'],'phases':j['rank0_phases_ms']}))"
done
