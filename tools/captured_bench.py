"""Eager render loop vs CUDA-graph replay (grafx_b200.render.CapturedRender) on a small mixing-console plan,
where issuing the launches costs more than running them.  usage: python tools/captured_bench.py [tracks] [B] [L]"""
import sys, time
import torch
sys.path.insert(0, ".")
import grafx_b200.processors as P
from grafx_b200.render import CapturedRender, mixing_console_plan, render_grafx

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
torch.manual_seed(0)
procs = {"eq": P.ParametricEqualizer().cuda(), "compressor": P.Compressor().cuda(), "reverb": P.STFTMaskedNoiseReverb(ir_len=24000).cuda()}
rd = mixing_console_plan(T, ["eq", "compressor", "reverb"])
x = torch.randn(B, T, 2, L, device="cuda")
prm = {k: {n: 0.5 * torch.randn(T, *((v,) if isinstance(v, int) else v), device="cuda") for n, v in p.parameter_size().items()}
       for k, p in procs.items()}
cap = CapturedRender(procs, x, prm, rd)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


e = timed(lambda: render_grafx(procs, x, prm, rd))
c = timed(lambda: cap(x, prm))
print(f"tracks={T} B={B} L={L}: eager {e:.3f} ms/render, captured {c:.3f} ms/render ({e / c:.2f}x)")
