#!/bin/bash
# after the empty-list prefetch fix: tests, racecheck on the tests that exercise empty lists / softened cells
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02ah_pytest_gpu.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge or random_lists or raw_particle or device_walk_bucket or clustered_box" > gpurun_out/r02ah_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02ah_sanitizer_racecheck.log; tail -4 gpurun_out/r02ah_sanitizer_racecheck.log
timeout 600 python tools/step_probe.py --n 16777216 --steps 3 > gpurun_out/r02ah_probe_256.json 2> /dev/null
python -c "
import json; j=json.load(open('gpurun_out/r02ah_probe_256.json')); r=j['resident']; print('256^3 resident', round(r['ms_per_step'],3), r['rank_phases_ms'])"
