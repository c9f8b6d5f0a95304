#!/bin/bash
# float filter in walk_node_fast: tests (lists bit-exact), pair counts at 256^3 and clustered, walk times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02aj_pytest_gpu.log
probe() { # name, n, kind, env...
  local name=$1; local n=$2; local kind=$3; shift; shift; shift
  env "$@" timeout 600 python tools/step_probe.py --n $n --kind $kind --steps 3 > gpurun_out/r02aj_probe_${n}_${kind}_$name.json 2> gpurun_out/r02aj_probe_${n}_${kind}_$name.err
  tail -2 gpurun_out/r02aj_probe_${n}_${kind}_$name.err
  python -c "
import json; j=json.load(open('gpurun_out/r02aj_probe_${n}_${kind}_$name.json')); r=j['resident']; print('$name $n $kind resident', round(r['ms_per_step'],3), r['rank_phases_ms'], 'pairs', j['pc_pairs'], j['pp_pairs'])"
}
probe filter 16777216 uniform X=1
probe filter 4194304 clustered X=1
probe filter 16777216 clustered X=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02aj_launches_step_256.csv \
  python tools/step_probe.py --n 16777216 --steps 1 > gpurun_out/r02aj_launches.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02aj_launches_step_256.csv')))
for i,r in enumerate(rows):
    if r and r[0]=='ID': h=i;break
hdr=rows[h]; seq=[]
for r in rows[h+1:]:
    if len(r)<len(hdr): continue
    d=dict(zip(hdr,r)); seq.append((d['Kernel Name'][:30], float(d['Metric Value'].replace(',',''))/1e3))
start=[i for i,s in enumerate(seq) if 'tree_keys' in s[0]]
b=seq[start[-1]:]
w=[x[1] for x in b if 'walk_level' in x[0]]
print('walk levels sum', round(sum(w),1), [round(x) for x in w[14:]])
print('pack', [round(x[1]) for x in b if 'walk_pack' in x[0]])
PY
