#!/bin/bash
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline --large-n 0 > gpurun_out/bench_n2_r01i.json 2> gpurun_out/bench_n2_r01i.err
tail -3 gpurun_out/bench_n2_r01i.err; python -c "
import json;j=json.load(open('gpurun_out/bench_n2_r01i.json'));print(json.dumps({'ms':j['ms_per_step'],'value':j['value'],'e2e':j['e2e']}))"
