#!/bin/bash
# two ranks: headline bench (weak scaling) + the 4 M box shared by both (strong), then the multistep clustered box
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2_r01i.json 2> gpurun_out/bench_n2_r01i.err
tail -2 gpurun_out/bench_n2_r01i.err; python -c "
import json;j=json.load(open('gpurun_out/bench_n2_r01i.json'));print(json.dumps({'ms':j['ms_per_step'],'value':j['value'],'e2e':j['e2e']['ms_per_step'],'large':{k:j['large_box'][k] for k in ('ms_per_step','rank0_phases_ms','finite')}}))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 --large-kind clustered --large-active-rung 2 > gpurun_out/bench_n2_multistep_r01i.json 2> gpurun_out/bench_n2_multistep_r01i.err
tail -2 gpurun_out/bench_n2_multistep_r01i.err; python -c "
import json;j=json.load(open('gpurun_out/bench_n2_multistep_r01i.json'));print(json.dumps(j['large_box']))"
