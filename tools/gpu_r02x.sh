#!/bin/bash
# eight GPUs: the bench line at N=8 (256^3 shared) with the clustered 512^3 target block; then N=4
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02x_bench_n8.json 2> gpurun_out/r02x_bench_n8.err
tail -4 gpurun_out/r02x_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 \
  bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02x_bench_n4.json 2> gpurun_out/r02x_bench_n4.err
python -c "
import json
for f in ('gpurun_out/r02x_bench_n8.json','gpurun_out/r02x_bench_n4.json'):
    j=json.load(open(f))
    print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e']['ms_per_step'], j['phases_ms_rank0'], j['roofline']['frac'], j['parity']['median_da_over_a'], j['parity_vs_n1'])
    if 'target' in j:
        t=j['target']; print({k:t.get(k) for k in ('force_step_ms','interactions_per_s','e2e_ms','rank0_phases_ms','rank0_hbm_in_use_gb','error')}, t.get('parity',{}).get('median_da_over_a'))
"
