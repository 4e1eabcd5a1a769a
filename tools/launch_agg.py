import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith(chr(34))))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = r[ki].split("(")[0]; v = float(r[vi].replace(",", ""))
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v / 1e3 if r[ui] == "ns" else v
for n, (c, t) in agg.items():
    if "walk" in n or "scan" in n: print(f"{n:40s} {c:5d} launches {t/1e3:9.3f} ms")
