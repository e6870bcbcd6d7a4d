"""Event timings of the fused dynamics kernel per CTA size of the scan variant (gfx_dynamics_set_tuning)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.processors as P
from grafx_b200 import _cabi


def timeit(fn, warm=3, it=15):
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
comp, gate = P.Compressor().cuda(), P.NoiseGate().cuda()
chain = P.SerialChain({"comp": comp, "gate": gate}).cuda()
shapes = [("cfg4 chain 1024x1x65536", 1024, 1, 65536, True), ("cfg5 compressor 512x2x131072", 512, 2, 131072, False),
          ("compressor 64x2x131072", 64, 2, 131072, False), ("compressor 8x2x131072", 8, 2, 131072, False)]
for name, B, C, L, is_chain in shapes:
    x = torch.randn(B, C, L, device="cuda")
    pc = {k: torch.randn(B, v, device="cuda") for k, v in comp.parameter_size().items()}
    pg = {k: torch.randn(B, v, device="cuda") for k, v in gate.parameter_size().items()}
    res = []
    for nt in (0, 128, 256):
        assert L_.gfx_dynamics_set_tuning(nt) == 0
        fn = (lambda: chain(x, comp=pc, gate=pg)) if is_chain else (lambda: comp(x, **pc))
        res.append(f"nt={nt}: {timeit(fn):.4f} ms")
    print(f"{name:32s} " + " | ".join(res), flush=True)
L_.gfx_dynamics_set_tuning(0)
