#!/bin/bash
# per-kernel times of one RawParticleStep (4 M uniform) under ncu
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_rawstep.csv \
  python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --large-n 4194304 > /dev/null 2>&1
python - <<'P'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_rawstep.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); vi=H.index('Metric Value'); ui=H.index('Metric Unit')
seq=[(r[ki], float(r[vi].replace(',',''))*{'ns':1e-3,'us':1,'ms':1e3}.get(r[ui],1e-3)) for r in rows[hdr+1:] if len(r)>vi and r[vi]]
idx=[i for i,(k,_) in enumerate(seq) if 'tree_keys' in k]
last=seq[idx[-1]:]
agg={}
for k,v in last:
    k=k.split('(')[0][-40:]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:12]: print(f"{k:42s} {n:3d} {v:9.1f}")
print([round(v) for k,v in last if 'walk_level' in k])
P
