"""Top stall sites of one kernel from `ncu -i rep --page source --csv` output (development aid).
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > f.csv; python tools/ncu_src.py f.csv [ntop]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
idx = {n: i for i, n in enumerate(h)}
data = []
for r in rows[hi + 1:]:
    if len(r) != len(h):
        continue
    try:
        int(r[idx["# Samples"]] or 0)
    except ValueError:
        continue
    data.append(r)
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print("kernel:", rows[0][1][:100] if rows[0] else "", "instructions", len(data), "samples", tot)
stalls = [n for n in h if n.startswith("stall_")]
agg = {n: sum(int(r[idx[n]] or 0) for r in data) for n in stalls}
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]] or 0))[:ntop]:
    s = int(r[idx["# Samples"]])
    st = sorted(((int(r[idx[n]] or 0), n[6:]) for n in stalls), reverse=True)[:3]
    print("%6d %5.1f%% %s  %-70s %s" % (s, 100.0 * s / max(tot, 1), r[idx["Address"]][-5:], r[idx["Source"]][:70], [x for x in st if x[0]]))
