"""Scaling probe of the ballistics dynamics path: batch and length sweeps (CUDA events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.processors as P

comp = P.Compressor(energy_smoother="ballistics").cuda()
cfgs = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(1024, 65536)]
for B, L in cfgs:
    x = torch.randn(B, 1, L, device="cuda")
    prm = {k: torch.randn(B, v, device="cuda") for k, v in comp.parameter_size().items()}
    torch.cuda.synchronize()
    print(f"B={B} L={L} start", flush=True)
    for i in range(3):
        t0 = time.perf_counter()
        comp(x, **prm)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
    print(f"B={B:5d} L={L:6d}: {ms:7.3f} ms   {ms*1e6/L:6.1f} ns/sample-step  ({ms*1e-3*1.965e9/L:5.1f} cycles)", flush=True)
