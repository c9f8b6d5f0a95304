#!/bin/bash
# tools/gpu_nN.sh N : the default bench at N ranks (weak scaling of the headline step + the shared 4 M box)
N=$1
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n${N}_r01i.json 2> gpurun_out/bench_n${N}_r01i.err
tail -2 gpurun_out/bench_n${N}_r01i.err; python -c "
import json;j=json.load(open('gpurun_out/bench_n${N}_r01i.json'));print(json.dumps({'n':j['n_gpus'],'ms':j['ms_per_step'],'value':j['value'],'e2e_ms':j['e2e']['ms_per_step'],'e2e':j['e2e']['value'],'large':{k:j['large_box'][k] for k in ('ms_per_step','interactions_per_s','rank0_phases_ms','finite')}}))"
