"""Prints the per-workload table of a set of bench.py JSON lines (profiles/r01_bench_*.json by default).
usage: python tools/roofline_report.py [files...]"""
import glob, json, os, sys

files = sys.argv[1:] or sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "profiles", "r01_bench_*.json")))
print("| file | workload | ms/step | samples/s | roofline (achieved / peak GB/s = frac) | kernel ms | traffic MB | e2e samples/s (ms) | launches | SM MHz |")
print("|---|---|---|---|---|---|---|---|---|---|")
for f in files:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f"| {os.path.basename(f)} | unreadable: {e} |")
        continue
    r, e, c = d.get("roofline") or {}, d.get("e2e") or {}, d.get("clocks") or {}
    wl = (d.get("config") or {}).get("workload", "")[:60]
    roof = f"{r.get('achieved', 0):.0f} / {r.get('peak', 0):.0f} = {r.get('frac', 0):.3f}" if r else "-"
    traffic = f"{r['traffic'] / 1e6:.0f}" if r.get("traffic") else "-"
    kms = f"{r['kernel_ms']:.3f}" if r.get("kernel_ms") else "-"
    e2e = f"{e.get('value', 0):.3e} ({e.get('ms_per_step', 0) or 0:.2f})" if e else "-"
    tag = "reference arm (CPU)" if d.get("impl") == "reference" else wl
    print(f"| {os.path.basename(f)} | {tag} | {d['ms_per_step']:.4f} | {d['value']:.3e} | {roof} | {kms} | {traffic} | {e2e} | "
          f"{d.get('gpu_launches', '-')} | {c.get('sm_mhz', '-')} |")
