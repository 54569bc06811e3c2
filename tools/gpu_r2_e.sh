#!/bin/bash
# round 2, trip E: suite, fused projection / anchor kernel timing after the instruction diet, stream-priority A/B, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_e.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_e.log
grep -E "^E  |passed|failed|exit|^FAILED" gpurun_out/pytest_e.log | cut -c1-250 | tail -30
timeout 300 python tools/time_proj.py > gpurun_out/time_proj2.txt 2>&1; cat gpurun_out/time_proj2.txt
timeout 300 python tools/time_attractor.py > gpurun_out/time_attractor2.txt 2>&1; cat gpurun_out/time_attractor2.txt
timeout 600 python tools/ab_switch.py GROUP_PRIORITIES=0 GROUP_PRIORITIES=1 USE_FUSED_PROJ_ANCHOR=0 > gpurun_out/ab_prio.txt 2>&1; cat gpurun_out/ab_prio.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_proj_launches2.csv python tools/time_proj.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/ncu_proj_launches2.csv')) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki][:60]].append(float(r[vi].replace(',', '')))
    except Exception: pass
for k, v in d.items():
    if 'danet' in k: print('%-62s n %3d  min %9.1f median %9.1f max %9.1f us' % (k, len(v), min(v)/1e3, sorted(v)[len(v)//2] / 1e3, max(v)/1e3))
PY
