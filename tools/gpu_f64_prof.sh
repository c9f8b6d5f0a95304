#!/bin/bash
# FP64 build: bench line + full ncu capture of the p-c / p-p / Ewald kernels
TAG=${1:-f64}
mkdir -p gpurun_out
timeout 600 python bench.py --double --steps 100 --warmup 5 --large-n 0 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cell_list|part_list|ewald_kernel" -s 9 -c 3 -f -o gpurun_out/prof_$TAG \
  python bench.py --double --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --large-n 0 > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
