import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    name = r[ki].split("(")[0].split("::")[-1][:60]
    agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
for k, v in agg.items():
    v2 = v[len(v) // 2:]  # second half: steady state
    print("%-60s n=%4d  mean %9.1f us  min %9.1f" % (k, len(v), sum(v2) / len(v2) / 1e3, min(v2) / 1e3))
