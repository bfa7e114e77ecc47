#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_qtc_e.csv \
    python tools/bench_configs.py c2a c2b c3 --small > gpurun_out/qtc_under_ncu_e.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_qtc_e.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]; kn = h.index('Kernel Name'); mv = h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    name = r[kn][:90]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, t) in agg.items():
    print(f"{t/1e6:10.3f} ms {c:5d}  {k}")
PY
