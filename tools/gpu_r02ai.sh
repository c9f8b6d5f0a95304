#!/bin/bash
# two GPUs: multistep step (active rung 2) and a plain step of the clustered 4 M box through the in-library path
mkdir -p gpurun_out
for rung in 0 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$rung \
  tools/step_probe.py --particles 4194304 --kind clustered --active-rung $rung --steps 3 > gpurun_out/r02ai_probe_4M_clustered_rung${rung}_n2.json 2> gpurun_out/r02ai_probe_rung${rung}.err
tail -2 gpurun_out/r02ai_probe_rung${rung}.err
python -c "
import json
s=open('gpurun_out/r02ai_probe_4M_clustered_rung${rung}_n2.json').read(); j=json.loads(s[s.index('{'):])
for m in ('e2e','resident'): print('rung $rung', m, round(j[m]['ms_per_step'],3), j[m]['rank_phases_ms'], j['pc_pairs'], j['pp_pairs'])"
done
