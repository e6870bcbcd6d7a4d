"""Runs one BASELINE workload a few times eagerly (for ncu launch lists / captures).  usage: run_workload_once.py <cfg> [renders]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
name = sys.argv[1]
with torch.no_grad():
    if name == "cfg5":
        from grafx_b200.render import render_grafx
        B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
        wl = bench.GraphWorkload()
        wl.build(torch.device("cuda"))
        prm = bench.tree_to(wl.host_params(), "cuda")
        x = torch.randn(B, wl.TRACKS, wl.C, wl.L, device="cuda")
        for _ in range(3):
            y = render_grafx(wl.procs, x, prm, wl.plan, parameters_grad=False)[0]
    else:
        wl = bench.Workload(name)
        wl.build("cuda")
        x, prm = wl.host_inputs()
        x = x.cuda(); prm = bench.tree_to(prm, "cuda")
        for _ in range(3):
            y = wl.forward(x, prm)
torch.cuda.synchronize()
