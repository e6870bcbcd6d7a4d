"""cfg4b (Compressor -> NoiseGate with ballistics) per chunk-size target of the speculative kernel, and other shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.processors as P
from grafx_b200 import _cabi


def timeit(fn, warm=3, it=11):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(it)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


L_ = _cabi.lib()
torch.manual_seed(0)
comp, gate = P.Compressor(energy_smoother="ballistics").cuda(), P.NoiseGate(energy_smoother="ballistics").cuda()
chain = P.SerialChain({"comp": comp, "gate": gate}).cuda()
for name, B, C, L in (("cfg4b 1024x1x65536", 1024, 1, 65536), ("512x2x131072", 512, 2, 131072), ("32x2x131072", 32, 2, 131072)):
    x = torch.randn(B, C, L, device="cuda")
    pc = {k: torch.randn(B, v, device="cuda") for k, v in comp.parameter_size().items()}
    pg = {k: torch.randn(B, v, device="cuda") for k, v in gate.parameter_size().items()}
    res = []
    L_.gfx_dynamics_set_ballistics_mode(0)
    res.append(f"walk {timeit(lambda: chain(x, comp=pc, gate=pg), 1, 3):.3f}")
    L_.gfx_dynamics_set_ballistics_mode(1)
    for tps in (384, 768, 1536, 3072, 6144):
        L_.gfx_dynamics_set_ballistics_mode(tps)
        res.append(f"tps{tps} {timeit(lambda: chain(x, comp=pc, gate=pg)):.3f}")
    print(f"{name:22s} " + " | ".join(res), flush=True)
