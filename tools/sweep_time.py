"""Long-filter convolution per sweep size (spectra workspace per sweep): time and, under ncu, DRAM traffic."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.functional as F_
from grafx_b200 import _cabi


def timeit(fn, warm=2, it=7):
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
x = torch.randn(512, 2, 131072, device="cuda")
h = torch.randn(512, 2, 96000, device="cuda") / 300
h32 = torch.randn(32, 2, 96000, device="cuda") / 300
only = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for mb in ([only] if only else [3072, 1536, 768]):
    L_.gfx_fir_set_sweep_mb(mb)
    if only:
        F_.fir_conv(x, h); torch.cuda.synchronize()
    else:
        a = timeit(lambda: F_.fir_conv(x, h))
        b = timeit(lambda: F_.fir_conv(x, h32, h_repeat=16))
        g = torch.cuda.CUDAGraph()
        y = F_.fir_conv(x, h)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            y = F_.fir_conv(x, h)
        c = timeit(g.replay)
        print(f"sweep {mb:5d} MiB: per-item filters {a:.3f} ms (graph replay {c:.3f}) | shared filters {b:.3f} ms", flush=True)
