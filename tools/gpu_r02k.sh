#!/bin/bash
# bucket pipeline (2-ahead grab, raw meta), packed reduction with RED/REDUX, target-pair prefetch, path-table emit_fill
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02k_pytest_gpu.log
probe() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python tools/step_probe.py --n 16777216 --steps 3 > gpurun_out/r02k_probe_256_$name.json 2> gpurun_out/r02k_probe_256_$name.err
  python -c "
import json; j=json.load(open('gpurun_out/r02k_probe_256_$name.json')); r=j['resident']; print('$name 256^3 resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pc', round(r['rank_pc_ms'],3), 'pp', round(r['rank_pp_ms'],3), 'ew', round(r['rank_ewald_ms'],3))"
}
probe new X=1
probe base CB200_LIB=changa_b200/variants/base.so
probe nopf CB200_LIB=changa_b200/variants/nopf.so
probe climb CB200_EMIT_CLIMB=1
timeout 300 python tools/step_probe.py --n 4194304 --steps 5 > gpurun_out/r02k_probe_4M.json 2> gpurun_out/r02k_probe_4M.err
python -c "
import json; j=json.load(open('gpurun_out/r02k_probe_4M.json')); r=j['resident']; print('4M resident', round(r['ms_per_step'],3), r['rank_phases_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02k_launches_step_256.csv \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02k_launches.log 2>&1
ls -la gpurun_out | tail -3
