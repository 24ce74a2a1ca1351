#!/bin/bash
# per-function instruction counts of tile_apply on config 3 (first / last / min / max)
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tile_apply -c 16 --csv --log-file gpurun_out/c3_apply.csv python bench.py --workload c3 --others none --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/c3_apply.csv")))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
h=rows[hi]; ik=h.index("Kernel Name"); im=h.index("Metric Name"); iv=h.index("Metric Value"); iid=h.index("ID")
d={}
for r in rows[hi+1:]:
    if len(r)<=iv: continue
    d.setdefault((r[iid], r[ik][:60]),{})[r[im]]=r[iv]
for k,v in d.items(): print(k, v)
PY
