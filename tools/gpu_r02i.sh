#!/bin/bash
# emit_fill with staged particle runs, p-p park split, moments at 8 CTAs/SM, FP64 p-c <6,2,pair> default; sanitizer logs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02i_pytest_gpu.log
timeout 300 python tools/step_probe.py --n 4194304 --steps 5 > gpurun_out/r02i_probe_4M.json 2> gpurun_out/r02i_probe_4M.err
tail -2 gpurun_out/r02i_probe_4M.err
python -c "
import json; j=json.load(open('gpurun_out/r02i_probe_4M.json')); r=j['resident']; print('4M resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pp pairs/s %.3e' % r['rank_pp_pairs_per_s'])"
timeout 600 python tools/step_probe.py --n 16777216 --steps 3 > gpurun_out/r02i_probe_256.json 2> gpurun_out/r02i_probe_256.err
python -c "
import json; j=json.load(open('gpurun_out/r02i_probe_256.json')); r=j['resident']; print('256^3 resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pp pairs/s %.3e' % r['rank_pp_pairs_per_s'])"
timeout 300 python tools/resident_probe.py --workload cube300 --double --steps 30 > gpurun_out/r02i_resident_cube300_f64.json 2>/dev/null; cat gpurun_out/r02i_resident_cube300_f64.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02i_launches_step_256.csv \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02i_launches.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge or random_lists or raw_particle or clustered_box or multistep or device_walk or device_tree" > gpurun_out/r02i_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02i_sanitizer_memcheck.log; tail -4 gpurun_out/r02i_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge or random_lists or raw_particle or device_walk_bucket" > gpurun_out/r02i_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02i_sanitizer_racecheck.log; tail -4 gpurun_out/r02i_sanitizer_racecheck.log
ls -la gpurun_out | tail -3
