"""Aggregates an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum[,gpu__time_duration.sum] --csv` log by kernel:
DRAM bytes read / written per repeat of the workload.  usage: python tools/traffic_agg.py <log.csv> [n_repeats] [scale]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
rep = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
agg = collections.OrderedDict()
for r in rows[1:]:
    if "dram__bytes" not in r[mi]:
        continue
    try:
        v = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[ki])[:70]
    a = agg.setdefault(name, [0.0, 0.0])
    a[0 if "read" in r[mi] else 1] += v
tr = tw = 0.0
for k, (rd, wr) in agg.items():
    print(f"{rd / rep / 1e6:10.1f} MB read  {wr / rep / 1e6:10.1f} MB written  {k}")
    tr += rd
    tw += wr
tot = (tr + tw) / rep
print(f"total per repeat: {tot / 1e9:.3f} GB  (x{scale:g} = {tot * scale / 1e9:.2f} GB, {int(tot * scale)} bytes)")
