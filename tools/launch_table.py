"""Print a per-launch table from an ncu --csv launch list (gpu__time_duration etc.)."""
import csv, sys
from collections import OrderedDict
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
d = OrderedDict()
for x in rows:
    d.setdefault(x['ID'], {'name': x['Kernel Name'][:46], 'grid': x['Grid Size']})[x['Metric Name']] = float(x['Metric Value'].replace(',', ''))
tot = sum(v['gpu__time_duration.sum'] for v in d.values())
for k, v in d.items():
    extra = ""
    if 'dram__bytes_read.sum' in v:
        extra = f" rd {v['dram__bytes_read.sum']/1e6:7.1f} MB wr {v['dram__bytes_write.sum']/1e6:7.1f} MB tensor {v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):5.1f}%"
    print(f"{v['name']:46s} {v['grid']:>14s} {v['gpu__time_duration.sum']/1e3:8.1f} us {100*v['gpu__time_duration.sum']/tot:5.1f}%{extra}")
print(f"total {tot/1e3:.1f} us")
