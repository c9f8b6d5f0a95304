#!/bin/bash
# what one rank of 8 does at 256^3, emulated on one GPU (CB200_EMULATE_RANK): phases and launch list
mkdir -p gpurun_out
for r in 0 3; do
CB200_EMULATE_RANK=$r/8 timeout 600 python tools/step_probe.py --n 16777216 --steps 3 > gpurun_out/r02s_probe_256_rank${r}of8.json 2> gpurun_out/r02s_probe_256_rank${r}of8.err
python -c "
import json; j=json.load(open('gpurun_out/r02s_probe_256_rank${r}of8.json')); r=j['resident']; print('rank $r/8 resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pairs', j['pc_pairs'], j['pp_pairs'])"
done
CB200_EMULATE_RANK=3/8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02s_launches_rank3of8.csv \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02s_launches.log 2>&1
CB200_EMULATE_RANK=3/8 timeout 600 nsys --version > /dev/null 2>&1 || true
