#!/bin/bash
# walk_node_fast (shared-memory-only lists, leaf specialisation, lean box distance) vs walk_node_general
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02m_pytest_gpu.log
probe() { # name, n, env...
  local name=$1; local n=$2; shift; shift
  env "$@" timeout 600 python tools/step_probe.py --n $n --steps 3 > gpurun_out/r02m_probe_${n}_$name.json 2> gpurun_out/r02m_probe_${n}_$name.err
  tail -2 gpurun_out/r02m_probe_${n}_$name.err
  python -c "
import json; j=json.load(open('gpurun_out/r02m_probe_${n}_$name.json')); r=j['resident']; print('$name $n resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pairs', j['pc_pairs'], j['pp_pairs'])"
}
probe fast 16777216 X=1
probe general 16777216 CB200_WALK_GENERAL=1
probe fast 4194304 X=1
timeout 300 python tools/step_probe.py --n 4194304 --kind clustered --steps 3 > gpurun_out/r02m_probe_4Mclu_fast.json 2> gpurun_out/r02m_probe_4Mclu_fast.err
CB200_WALK_GENERAL=1 timeout 300 python tools/step_probe.py --n 4194304 --kind clustered --steps 3 > gpurun_out/r02m_probe_4Mclu_general.json 2> gpurun_out/r02m_probe_4Mclu_general.err
python -c "
import json
for k in ('fast','general'):
    j=json.load(open('gpurun_out/r02m_probe_4Mclu_%s.json'%k)); r=j['resident']; print(k,'4M clustered', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pairs', j['pc_pairs'], j['pp_pairs'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02m_launches_step_256.csv \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02m_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"walk_level" -s 102 -c 4 -f -o gpurun_out/r02m_prof_walk_4M \
  python tools/step_probe.py --n 4194304 --steps 1 > gpurun_out/r02m_prof_walk.log 2>&1
ls -la gpurun_out | tail -3
