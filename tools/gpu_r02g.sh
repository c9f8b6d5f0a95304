#!/bin/bash
# eight GPUs: the bench line at N=8 (256^3 shared) with the clustered 512^3 target block
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02g_bench_n8.json 2> gpurun_out/r02g_bench_n8.err
tail -12 gpurun_out/r02g_bench_n8.err; cat gpurun_out/r02g_bench_n8.json
