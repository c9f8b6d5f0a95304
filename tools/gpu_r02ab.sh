#!/bin/bash
# bucket draw by inline PTX (no compiler-made warp aggregation), reset check at the broadcast
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02ab_pytest_gpu.log
for n in 16777216 4194304; do
timeout 600 python tools/step_probe.py --n $n --steps 3 > gpurun_out/r02ab_probe_$n.json 2> gpurun_out/r02ab_probe_$n.err
tail -2 gpurun_out/r02ab_probe_$n.err
python -c "
import json; j=json.load(open('gpurun_out/r02ab_probe_$n.json')); r=j['resident']; print('$n resident', round(r['ms_per_step'],3), r['rank_phases_ms'])"
done
timeout 300 python tools/resident_probe.py --workload cube300 2>/dev/null | tail -1 | cut -c1-600
