#!/bin/bash
# the new bench line on BASELINE config 3 (uniform 256^3), one GPU: smoke, parity tests, both arms
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r02d_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02d_pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
tail -5 gpurun_out/r02d_bench.err; cat gpurun_out/r02d_bench.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02d_bench_reference.json 2> gpurun_out/r02d_bench_reference.err
tail -3 gpurun_out/r02d_bench_reference.err; cat gpurun_out/r02d_bench_reference.json
nproc; free -g | head -2
