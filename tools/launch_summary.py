"""Aggregate an ncu launch list (gpu__time_duration) by kernel: python tools/launch_summary.py file.csv [skip_launches]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; data = rows[hi + 1:]
ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    n = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n[:60]:60s} launches {a[0]:5d}  total {a[1]/1e3:10.1f} us  {100*a[1]/tot:5.1f}%")
print("total us", tot / 1e3)
