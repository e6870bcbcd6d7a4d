"""Round-2 summary table from one default bench.py line (main workload + per_config records) and the reference arm.
usage: python tools/roofline_report_r02.py [bench_final.json] [bench_reference.json] > profiles/r02_summary.md"""
import json, os, sys

here = os.path.join(os.path.dirname(__file__), "..", "profiles")
f = sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "r02_bench_final.json")
g = sys.argv[2] if len(sys.argv) > 2 else os.path.join(here, "r02_bench_reference.json")
d = json.loads(open(f).read().strip().splitlines()[-1])
ref = json.loads(open(g).read().strip().splitlines()[-1])
print("| workload | ms/step | samples/s | HBM roofline (achieved / peak GB/s = frac) | fp32-FMA frac | kernel ms | DRAM traffic MB (algorithmic MB) | CPU arm samples/s (threads) | GPU / CPU |")
print("|---|---|---|---|---|---|---|---|---|")


def row(name, rec, cpu):
    r = rec["roofline"]
    tr = f"{r['traffic'] / 1e6:.0f}" if r.get("traffic") else "-"
    print(f"| {name} | {rec['ms_per_step']:.4f} | {rec['value']:.3e} | {r['achieved']:.0f} / {r['peak']:.0f} = {r['frac']:.3f} | "
          f"{r.get('fp32_fma_frac', 0):.3f} | {r['kernel_ms']:.4f} | {tr} ({r['algorithmic_bytes'] / 1e6:.0f}) | "
          f"{cpu['value']:.3e} ({cpu['cores']}) | {rec['value'] / cpu['value']:.0f}x |")


row("cfg5 (main): " + d["config"]["workload"][:70], d, d["cpu_baseline"])
for k, v in d["per_config"].items():
    row(f"{k}: " + v["workload"][:70], v, v["cpu_baseline"])
e = d["e2e"]
print()
print(f"End to end (host buffers, copies inside), config 5 on one GPU: {e['value']:.3e} samples/s, {e['ms_per_step']:.1f} ms per step "
      f"({e['h2d_bytes_per_step'] / 1e9:.2f} GB in, {e['d2h_bytes_per_step'] / 1e6:.0f} MB out; host-to-device ceiling {e['ceiling_ms_per_step']:.1f} ms = "
      f"{e['ceiling_gbs_per_rank']:.1f} GB/s).  Reference arm (`--impl reference`, same config object, {ref['cpu_baseline']['cores']} host threads): "
      f"{ref['value']:.3e} samples/s ({ref['ms_per_step']:.0f} ms per rendered graph).  Clocks: {d['clocks']['sm_mhz']} MHz median under load, "
      f"reasons {d['clocks']['reasons']}.")
