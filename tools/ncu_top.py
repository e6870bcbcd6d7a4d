"""Top stall lines of an `ncu --page source --csv` dump.  usage: ncu_top.py <rep> [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
isrc = hdr.index('Source'); ins = hdr.index('# Samples'); ie = hdr.index('Instructions Executed')
data = []
for i, r in enumerate(rows[2:]):
    try: data.append((int(r[ins]), i, r[isrc].strip()[:100], int(r[ie])))
    except Exception: pass
tot = sum(d[0] for d in data)
print('total samples', tot)
for s, i, src, e in sorted(data, reverse=True)[:n]:
    print(f"{s:6d} {100*s/tot:5.1f}%  #{i:5d} exec={e:9d}  {src}")
