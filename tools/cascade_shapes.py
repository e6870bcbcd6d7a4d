"""Device time of the biquad cascade over batch sizes / section counts (chain-latency-bound to HBM-bound)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.functional as F_

def timeit(fn, warm=3, it=9):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(it)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]

torch.manual_seed(0)
for B, C, L, K in ((1, 1, 131072, 5), (2, 2, 131072, 5), (8, 2, 131072, 5), (32, 2, 131072, 5), (256, 2, 131072, 1),
                   (256, 2, 131072, 5), (256, 2, 131072, 10), (64, 2, 131072, 31), (1024, 2, 16384, 5)):
    x = torch.randn(B, C, L, device="cuda")
    Bs = torch.randn(B, C, K, 3, device="cuda") * 0.1
    As = torch.randn(B, C, K, 3, device="cuda") * 0.1
    Bs[..., 0] += 1; As[..., 0] += 1
    ms = timeit(lambda: F_.biquad_cascade(x, Bs, As))
    print(f"B={B} C={C} L={L} K={K}: {ms*1e3:.1f} us")
