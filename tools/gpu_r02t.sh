#!/bin/bash
# per-level node ranges in the walk (one rank of 8 emulated), lazy node particles
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02t_pytest_gpu.log
for r in 0/1 0/8 3/8 7/8 1/2; do
tag=$(echo $r | tr / of)
CB200_EMULATE_RANK=$r timeout 600 python tools/step_probe.py --n 16777216 --steps 3 > gpurun_out/r02t_probe_256_rank${tag}.json 2> gpurun_out/r02t_probe_256_rank${tag}.err
tail -2 gpurun_out/r02t_probe_256_rank${tag}.err
python -c "
import json; j=json.load(open('gpurun_out/r02t_probe_256_rank${tag}.json')); r=j['resident']; print('rank $r resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pairs', j['pc_pairs'], j['pp_pairs'])"
done
