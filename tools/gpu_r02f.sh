#!/bin/bash
# walk reading the parent's undecided list in place; list kernels over the rank's range; configs 1/2/5
# kernel numbers; FP64 p-c variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02f_pytest_gpu.log
timeout 300 python tools/step_probe.py --n 4194304 --steps 5 > gpurun_out/r02f_probe_4M.json 2> gpurun_out/r02f_probe_4M.err
tail -2 gpurun_out/r02f_probe_4M.err
python -c "
import json; j=json.load(open('gpurun_out/r02f_probe_4M.json')); r=j['resident']; print('4M resident', r['ms_per_step'], r['rank_phases_ms'])"
for w in cube300 king; do
  timeout 300 python tools/resident_probe.py --workload $w --steps 50 > gpurun_out/r02f_resident_$w.json 2> gpurun_out/r02f_resident_$w.err
  tail -2 gpurun_out/r02f_resident_$w.err; cat gpurun_out/r02f_resident_$w.json
done
timeout 300 python tools/resident_probe.py --workload collapse --double --steps 50 > gpurun_out/r02f_resident_collapse_f64.json 2> gpurun_out/r02f_resident_collapse_f64.err
cat gpurun_out/r02f_resident_collapse_f64.json
bash tools/gpu_f64_ab.sh 0 4 5 6 2>&1 | tee gpurun_out/r02f_f64_ab.log
for v in 4 5 6; do CB200_PC64_VARIANT=$v timeout 300 python -m pytest tests -m gpu -q -k "double or collapse" 2>&1 | tail -2; done | tee gpurun_out/r02f_f64_variant_parity.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02f_launches_step_4M.csv \
  python tools/step_probe.py --n 4194304 --steps 1 > gpurun_out/r02f_launches.log 2>&1
ls -la gpurun_out | tail -3
