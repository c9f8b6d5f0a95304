#!/bin/bash
# moments kernel with staged exports; ncu --set full captures of the 256^3 step's top kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02j_pytest_gpu.log
timeout 300 python tools/step_probe.py --n 4194304 --steps 5 > gpurun_out/r02j_probe_4M.json 2> gpurun_out/r02j_probe_4M.err
tail -2 gpurun_out/r02j_probe_4M.err
python -c "
import json; j=json.load(open('gpurun_out/r02j_probe_4M.json')); r=j['resident']; print('4M resident', round(r['ms_per_step'],3), r['rank_phases_ms'])"
timeout 600 python tools/step_probe.py --n 16777216 --steps 3 > gpurun_out/r02j_probe_256.json 2> gpurun_out/r02j_probe_256.err
python -c "
import json; j=json.load(open('gpurun_out/r02j_probe_256.json')); r=j['resident']; print('256^3 resident', round(r['ms_per_step'],3), r['rank_phases_ms'])"
# one step = 4 warm (3 + resident) ... capture the LAST launch of each kernel family of a 256^3 run
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"cell_list_x2" -s 4 -c 1 -f -o gpurun_out/r02j_prof_pc_256 \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02j_prof_pc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"part_list_stream" -s 4 -c 1 -f -o gpurun_out/r02j_prof_pp_256 \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02j_prof_pp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"emit_fill|build_moments|ewald_slot" -s 60 -c 12 -f -o gpurun_out/r02j_prof_misc_4M \
  python tools/step_probe.py --n 4194304 --steps 1 > gpurun_out/r02j_prof_misc.log 2>&1
ls -la gpurun_out | tail -4
