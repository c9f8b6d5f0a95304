#!/bin/bash
# eight GPUs: the clustered 512^3 target through step_probe (5 e2e + 5 resident steps after 3 warm-up steps)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 \
  tools/step_probe.py --particles 134217728 --kind clustered --steps 5 > gpurun_out/r02aq_probe_512_clustered_n8.json 2> gpurun_out/r02aq_probe_512.err
tail -3 gpurun_out/r02aq_probe_512.err
python -c "
import json
s=open('gpurun_out/r02aq_probe_512_clustered_n8.json').read(); j=json.loads(s[s.index('{'):])
for m in ('e2e','resident'): print(m, round(j[m]['ms_per_step'],3), j[m]['rank_phases_ms'])
print(j['pc_pairs'], j['pp_pairs'])"
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -2
