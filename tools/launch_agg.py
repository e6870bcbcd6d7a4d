"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_agg.py <launches.csv> [n_repeats]   (times are per repeat)"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
rep = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    u = r[ui]
    v_us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", r[ki])[:90]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v_us
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t/rep:10.1f} us {100*t/tot:5.1f}%  n={n/rep:6.1f}  {k}")
print(f"{tot/rep:10.1f} us total per repeat")
