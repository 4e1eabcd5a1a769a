#!/bin/bash
# Round 2, call M: ncu launch list (gpu__time_duration per kernel, serialised) of ONE end-to-end C2 call at full scale.
set -u
O=gpurun_out
TAG=${1:-r2m}
mkdir -p $O
python - <<PY
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import workloads as W
from bench import data_dir
print(W.make_bam("c2", 1.0, data_dir())[0])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_e2e_c2.csv \
    python tools/e2e_ab.py --preset c2 --reps -1 base: > $O/${TAG}_launches.log 2>&1
tail -2 $O/${TAG}_launches.log
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("$O/${TAG}_launches_e2e_c2.csv") if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split("(")[0]; v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui] == "ns" else v
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
for n, (c, t) in agg.items(): print(f"{n:40s} {c:5d} launches {t/1e3:9.3f} ms")
PY
