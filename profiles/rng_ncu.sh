#!/bin/bash
# launch list of one device draw plan (64 slots x 4 classes of 250k rows): per-kernel device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_rng_launches.csv python profiles/rng_probe.py 250000 64 > gpurun_out/r2_rng_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_rng_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names = rows[hdr]
ki, vi = names.index("Kernel Name"), names.index("Metric Value")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 2:]:
    if len(r) > vi:
        k = r[ki].split("(")[0][-40:]
        tot[k][0] += 1
        tot[k][1] += float(r[vi].replace(",", "")) / 1e6
for k, (n, ms) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{k:42s} {n:5d} launches {ms:9.3f} ms  {ms/n*1e3:9.1f} us each")
PY
