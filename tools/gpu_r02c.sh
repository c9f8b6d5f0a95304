#!/bin/bash
# p-p streaming kernel, single code copy: parity, A/B at 4 M, ncu of p-p and of the leaf walk levels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02c_pytest_gpu.log
for v in 0 3 2; do
  CB200_PP_VARIANT=$v timeout 300 python tools/step_probe.py --n 4194304 --steps 5 > gpurun_out/r02c_probe_4M_pp$v.json 2> gpurun_out/r02c_probe_4M_pp$v.err
  tail -2 gpurun_out/r02c_probe_4M_pp$v.err
  python -c "
import json; j=json.load(open('gpurun_out/r02c_probe_4M_pp$v.json')); r=j['resident']; print('pp variant $v: pp_ms %.3f pairs/s %.3e  step %.2f ms' % (r['rank_pp_ms'], r['rank_pp_pairs_per_s'], r['ms_per_step']))"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"part_list" -s 3 -c 1 -f -o gpurun_out/r02c_prof_pp \
  python tools/step_probe.py --n 4194304 --steps 1 > gpurun_out/r02c_prof_pp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"walk_level" -s 60 -c 6 -f -o gpurun_out/r02c_prof_walk \
  python tools/step_probe.py --n 4194304 --steps 1 > gpurun_out/r02c_prof_walk.log 2>&1
ls -la gpurun_out | tail -3
