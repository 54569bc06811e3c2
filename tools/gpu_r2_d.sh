#!/bin/bash
# round 2, trip D: whole GPU suite, fused projection timing, ncu on the fused projection and the anchor kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_d.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_d.log
grep -E "^E  |passed|failed|exit|^FAILED" gpurun_out/pytest_d.log | cut -c1-300 | tail -30
timeout 300 python tools/time_proj.py > gpurun_out/time_proj.txt 2>&1; cat gpurun_out/time_proj.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_proj_launches.csv python tools/time_proj.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/ncu_proj_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki][:60]].append(float(r[vi].replace(',', '')))
    except Exception: pass
for k, v in d.items(): print('%-62s n %3d  median %9.1f us' % (k, len(v), sorted(v)[len(v)//2] / 1e3))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:proj_anchor_kernel -s 3 -c 1 -o gpurun_out/prof_proj_anchor python tools/time_proj.py > gpurun_out/ncu_full_proj.log 2>&1; tail -2 gpurun_out/ncu_full_proj.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:anchor2_mma -s 3 -c 1 -o gpurun_out/prof_anchor_mma python tools/time_attractor.py > gpurun_out/ncu_full_att.log 2>&1; tail -2 gpurun_out/ncu_full_att.log
ls -la gpurun_out/*.ncu-rep
