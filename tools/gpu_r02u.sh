#!/bin/bash
# the bench line on one GPU (default flags) and the reference arm
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r02u_bench_256_1gpu.json 2> gpurun_out/r02u_bench_256_1gpu.err ) 2>&1 | tail -3
tail -3 gpurun_out/r02u_bench_256_1gpu.err
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02u_bench_reference_arm.json 2> gpurun_out/r02u_bench_reference_arm.err ) 2>&1 | tail -3
python -c "
import json
j=json.load(open('gpurun_out/r02u_bench_256_1gpu.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e']['ms_per_step'], j['phases_ms_rank0'], j['roofline']['frac'], j['kernels']['pp_frac_of_fp32_peak'], j['parity']['median_da_over_a'], j['clocks'])
print(j.get('ref_cuda'))
print(j.get('cpu_baseline'))
r=json.load(open('gpurun_out/r02u_bench_reference_arm.json')); print(r['value'], r['ms_per_step'], r['cpu_baseline'])
"
