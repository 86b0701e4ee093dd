"""Per-instruction stall samples of one kernel out of an `ncu --set full --import-source on` report.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > k.csv; python tools/sass_stalls.py k.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ends = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
rows = rows[ends[0]:ends[1]]  # the report repeats the section per view: the first one is enough
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
num = lambda r, k: int(float(r[idx[k]] or 0))
tot = sum(num(r, "# Samples") for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s[6:]: sum(num(r, s) for r in data) for s in stalls}
print(rows[0][1][:100])
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
top = sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:top_n]
for i in sorted(top):
    r = data[i]
    st = {s[6:]: num(r, s) for s in stalls if num(r, s) > 0}
    st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
    print(f"{i:5d} {r[idx['Source']].strip()[:64]:64s} {num(r, '# Samples'):6d} {100.0 * num(r, '# Samples') / tot:5.1f}% exec {num(r, 'Instructions Executed'):8d} {st}")
