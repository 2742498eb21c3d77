#!/bin/bash
# 1 GPU: ncu launch list of the driver's bench command with the round's final library
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02at.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --parity-reads 20000 > gpurun_out/ncu_launches_r02at.log 2>&1
tail -1 gpurun_out/ncu_launches_r02at.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_r02at.csv")) if len(r) > 10]
hdr = rows[0]; k = hdr.index("Kernel Name"); v = hdr.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    try: ns = float(r[v].replace(",", ""))
    except ValueError: continue
    name = r[k].split("(")[0][:70]
    tot[name] += ns; cnt[name] += 1
for name, ns in tot.most_common(14): print("%9.3f ms %5d x  %s" % (ns / 1e6, cnt[name], name))
PY
