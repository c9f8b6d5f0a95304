#!/bin/bash
# two GPUs: the NCCL step test, then the bench line at N=2 (256^3 shared, strong scaling)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_step_multigpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02e_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02e_bench_n2.json 2> gpurun_out/r02e_bench_n2.err
tail -8 gpurun_out/r02e_bench_n2.err; cat gpurun_out/r02e_bench_n2.json
