#!/bin/bash
# eight GPUs: phases of the end-to-end step (h2d / finish) at 256^3
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
  tools/step_probe.py --particles 16777216 --steps 5 > gpurun_out/r02ac_probe_256_n8.json 2> gpurun_out/r02ac_probe_256_n8.err
tail -3 gpurun_out/r02ac_probe_256_n8.err
python -c "
import json; j=json.load(open('gpurun_out/r02ac_probe_256_n8.json'))
for m in ('e2e','resident'): print(m, round(j[m]['ms_per_step'],3), j[m]['rank_phases_ms'])"
nvidia-smi topo -m | head -14
numactl -H 2>/dev/null | head -6
