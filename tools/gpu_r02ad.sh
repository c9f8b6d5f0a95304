#!/bin/bash
# slab-wise result copy-out: tests, bench line at N=1 (e2e with SFC-order rows)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02ad_pytest_gpu.log
python bench.py --no-cpu-baseline --no-ref-cuda > gpurun_out/r02ad_bench_256_1gpu.json 2> gpurun_out/r02ad_bench_256_1gpu.err
tail -3 gpurun_out/r02ad_bench_256_1gpu.err
python -c "
import json
j=json.load(open('gpurun_out/r02ad_bench_256_1gpu.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['phases_ms_rank0'], j['roofline']['frac'], j['parity']['median_da_over_a'])
"
