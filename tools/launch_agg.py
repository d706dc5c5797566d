"""Aggregates an ncu launch list (gpu__time_duration.sum, --csv) per kernel.  usage: python tools/launch_agg.py list.csv [other.csv]"""
import csv, collections, sys

def agg(f):
    rows = list(csv.reader(open(f)))
    hdr, a = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            k = d["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")[:40]
            v = float(d["Metric Value"].replace(",", "")) / 1e3
            e = a.setdefault(k, [0, 0.0])
            e[0] += 1
            e[1] += v
    return a

A = agg(sys.argv[1])
B = agg(sys.argv[2]) if len(sys.argv) > 2 else None
keys = list(A.keys()) + ([k for k in B if k not in A] if B else [])
tot = [0.0, 0.0]
for k in sorted(keys, key=lambda k: -(A.get(k, [0, 0])[1])):
    a = A.get(k, [0, 0.0])
    line = f"{a[1]:10.1f} us x{a[0]:3d}"
    tot[0] += a[1]
    if B is not None:
        b = B.get(k, [0, 0.0])
        tot[1] += b[1]
        line += f"   | {b[1]:10.1f} us x{b[0]:3d}  {b[1] - a[1]:+9.1f}"
    print(line, " ", k)
print("total", [round(x, 1) for x in tot])
