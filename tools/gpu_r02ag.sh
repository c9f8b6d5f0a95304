#!/bin/bash
# final state: other configurations (resident probes), compute-sanitizer memcheck / racecheck over the list-kernel,
# device-tree, walk, emit and in-library step tests
mkdir -p gpurun_out
timeout 300 python tools/step_probe.py --n 16777216 --kind clustered --steps 3 > gpurun_out/r02ag_probe_256_clustered.json 2> gpurun_out/r02ag_probe_256_clustered.err
python -c "
import json; j=json.load(open('gpurun_out/r02ag_probe_256_clustered.json')); r=j['resident']; print('256^3 clustered resident', round(r['ms_per_step'],3), r['rank_phases_ms'])"
timeout 300 python tools/step_probe.py --n 4194304 --steps 5 > gpurun_out/r02ag_probe_4M.json 2> /dev/null
timeout 300 python tools/step_probe.py --n 4194304 --kind clustered --steps 5 > gpurun_out/r02ag_probe_4M_clustered.json 2> /dev/null
timeout 300 python tools/step_probe.py --n 4194304 --kind clustered --active-rung 2 --steps 5 > gpurun_out/r02ag_probe_4M_clustered_rung2.json 2> gpurun_out/r02ag_probe_4M_clustered_rung2.err
python -c "
import json
for f in ('4M','4M_clustered','4M_clustered_rung2'):
    j=json.load(open('gpurun_out/r02ag_probe_%s.json'%f)); r=j['resident']; print(f, round(r['ms_per_step'],3), r['rank_phases_ms'])"
timeout 300 python tools/resident_probe.py --workload cube300 --steps 30 > gpurun_out/r02ag_resident_cube300.json 2>/dev/null; cut -c1-400 gpurun_out/r02ag_resident_cube300.json
timeout 300 python tools/resident_probe.py --workload king --steps 30 > gpurun_out/r02ag_resident_king.json 2>/dev/null; cut -c1-300 gpurun_out/r02ag_resident_king.json
timeout 300 python tools/resident_probe.py --workload collapse --double --steps 30 > gpurun_out/r02ag_resident_collapse_f64.json 2>/dev/null; cut -c1-300 gpurun_out/r02ag_resident_collapse_f64.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge or random_lists or raw_particle or clustered_box or multistep or device_walk or device_tree or king_energy or reference_fixture" > gpurun_out/r02ag_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02ag_sanitizer_memcheck.log; tail -4 gpurun_out/r02ag_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge or random_lists or raw_particle or device_walk_bucket or clustered_box" > gpurun_out/r02ag_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02ag_sanitizer_racecheck.log; tail -4 gpurun_out/r02ag_sanitizer_racecheck.log
