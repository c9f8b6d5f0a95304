#!/bin/bash
# eight GPUs, final default state: the 256^3 bench line without the 512^3 block
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus 8 --steps 10 --warmup 3 --no-target > gpurun_out/r02ar_bench_n8.json 2> gpurun_out/r02ar_bench_n8.err
tail -2 gpurun_out/r02ar_bench_n8.err
python -c "
import json
j=json.load(open('gpurun_out/r02ar_bench_n8.json'))
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','hbm_in_use_gb_rank0')}, j['e2e']['ms_per_step'], j['phases_ms_rank0'], j['parity']['median_da_over_a'], j['parity_vs_n1']['bitwise_equal'])
"
