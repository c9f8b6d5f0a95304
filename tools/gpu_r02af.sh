#!/bin/bash
# two GPUs: locally essential moment build -- NCCL test (LET off / forced on at small size), bench line at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_multigpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02af_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 2 --steps 10 --warmup 3  > gpurun_out/r02af_bench_n2.json 2> gpurun_out/r02af_bench_n2.err
tail -4 gpurun_out/r02af_bench_n2.err
python -c "
import json
j=json.load(open('gpurun_out/r02af_bench_n2.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e']['ms_per_step'], j['phases_ms_rank0'], j['config']['moment_build'], j['parity']['median_da_over_a'], j['parity_vs_n1'])
"
