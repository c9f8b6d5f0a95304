#!/bin/bash
# walk with shared-memory list heads; moments kernel occupancy variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02h_pytest_gpu.log
for v in 0 3 1 2; do
  CB200_MOM_VARIANT=$v timeout 300 python tools/step_probe.py --n 4194304 --steps 5 > gpurun_out/r02h_probe_4M_mom$v.json 2> gpurun_out/r02h_probe_4M_mom$v.err
  tail -2 gpurun_out/r02h_probe_4M_mom$v.err
  python -c "
import json; j=json.load(open('gpurun_out/r02h_probe_4M_mom$v.json')); r=j['resident']; print('mom variant $v: 4M resident', round(r['ms_per_step'],3), r['rank_phases_ms'])"
done
timeout 600 python tools/step_probe.py --n 16777216 --steps 3 > gpurun_out/r02h_probe_256.json 2> gpurun_out/r02h_probe_256.err
python -c "
import json; j=json.load(open('gpurun_out/r02h_probe_256.json')); r=j['resident']; print('256^3 resident', round(r['ms_per_step'],3), r['rank_phases_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02h_launches_step_256.csv \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02h_launches.log 2>&1
ls -la gpurun_out | tail -3
