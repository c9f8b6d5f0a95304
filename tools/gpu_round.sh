#!/bin/bash
# one GPU box visit: parity tests, bench, ncu launch list, full ncu capture of the list kernels
# usage: tools/gpu_round.sh [tag] [quick]
TAG=${1:-run}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
if [ "$2" == "quick" ]; then exit 0; fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --large-n 0 > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cell_list -s 3 -c 1 -f -o gpurun_out/prof_cell_$TAG \
  python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --large-n 0 > gpurun_out/prof_cell.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"part_list|ewald_kernel" -s 6 -c 2 -f -o gpurun_out/prof_part_ewald_$TAG \
  python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --large-n 0 > gpurun_out/prof_pe.log 2>&1
ls -la gpurun_out | tail -5
