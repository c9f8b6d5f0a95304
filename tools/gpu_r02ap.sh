#!/bin/bash
# node arrays and walk pools sized from the last step (with retries): tests, HBM in use, step times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02ap_pytest_gpu.log
python bench.py --no-cpu-baseline --no-ref-cuda > gpurun_out/r02ap_bench_256_1gpu.json 2> gpurun_out/r02ap_bench_256_1gpu.err
tail -2 gpurun_out/r02ap_bench_256_1gpu.err
python -c "
import json
j=json.load(open('gpurun_out/r02ap_bench_256_1gpu.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','hbm_in_use_gb_rank0')}, j['e2e']['ms_per_step'], j['phases_ms_rank0'])
"
