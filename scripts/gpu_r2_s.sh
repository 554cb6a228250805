#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "bwd_weight or comb" 2>&1 | tail -n 2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2s_launches.csv python scripts/profile_step.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2s_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.OrderedDict(); tot=0
for r in rows[1:]:
    n=r[ki].split("(")[0].replace("void ","").replace("gte::","")[:50]; v=float(r[vi])/1e3
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
print("launches", len(rows)-1, "total us", round(tot,1))
for n,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{n:52s} {a[0]:2d} {a[1]:8.1f}")
PY
