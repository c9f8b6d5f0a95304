#!/bin/bash
# round 2, box visit b: parity (streaming p-p kernel, Ewald vs the reference kernel, ABI link test, the
# in-library step), p-p A/B, Ewald overlap A/B, 256^3, ncu of the p-p kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/r02b_pytest_gpu.log
for v in 0 3 2; do
  CB200_PP_VARIANT=$v timeout 300 python tools/step_probe.py --n 4194304 --steps 5 > gpurun_out/r02b_probe_4M_pp$v.json 2> gpurun_out/r02b_probe_4M_pp$v.err
  tail -2 gpurun_out/r02b_probe_4M_pp$v.err; cat gpurun_out/r02b_probe_4M_pp$v.json
done
timeout 300 python tools/step_probe.py --n 4194304 --steps 5 --no-overlap > gpurun_out/r02b_probe_4M_nooverlap.json 2> gpurun_out/r02b_probe_4M_nooverlap.err
cat gpurun_out/r02b_probe_4M_nooverlap.json
timeout 600 python tools/step_probe.py --n 16777216 --steps 3 > gpurun_out/r02b_probe_256.json 2> gpurun_out/r02b_probe_256.err
tail -2 gpurun_out/r02b_probe_256.err; cat gpurun_out/r02b_probe_256.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"part_list" -s 3 -c 1 -f -o gpurun_out/r02b_prof_pp \
  python tools/step_probe.py --n 4194304 --steps 1 > gpurun_out/r02b_prof_pp.log 2>&1
ls -la gpurun_out | tail -4
