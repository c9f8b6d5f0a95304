#!/bin/bash
# adaptive output slabs: tests + N=1 bench line (full) + reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02al_pytest_gpu.log
python bench.py > gpurun_out/r02al_bench_256_1gpu.json 2> gpurun_out/r02al_bench_256_1gpu.err
tail -3 gpurun_out/r02al_bench_256_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02al_bench_reference_arm.json 2> gpurun_out/r02al_bench_reference_arm.err
python -c "
import json
j=json.load(open('gpurun_out/r02al_bench_256_1gpu.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e']['ms_per_step'], j['phases_ms_rank0'], j['roofline']['frac'], j['kernels']['pp_frac_of_fp32_peak'], j['parity']['median_da_over_a'])
print(j['ref_cuda']['speedup'], j['cpu_baseline']['value'])
r=json.load(open('gpurun_out/r02al_bench_reference_arm.json')); print(r['value'], r['ms_per_step'])
"
