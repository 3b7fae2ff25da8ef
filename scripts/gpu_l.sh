#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/l_launches_c2.csv python scripts/run_configs.py c2 > gpurun_out/l_c2.log 2>&1
echo "rc=$? lines=$(wc -l < gpurun_out/l_launches_c2.csv)"; tail -1 gpurun_out/l_c2.log | cut -c1-300
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/l_launches_c2.csv')) if len(r)>5]
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[start+1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg[r[ki].split('(')[0][:48]][0]+=1; agg[r[ki].split('(')[0][:48]][1]+=v
tot=sum(v for _,v in agg.values())
for k,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:10]: print(f"{k:50s} {c:5d} {v/1e6:10.3f} ms {100*v/tot:5.1f}%")
PY
