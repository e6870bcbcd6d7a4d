"""Bus sum / source copy of one config-5 chunk (16 renders, 32 tracks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.functional as F_


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


buf = torch.randn(129, 16, 2, 131072, device="cuda")
x = torch.randn(16, 32, 2, 131072, device="cuda")
src = buf.narrow(0, 96, 32).transpose(0, 1)
out = buf.narrow(0, 128, 1).transpose(0, 1)
t = timeit(lambda: F_.node_sum(src, 1, None, 1, out=out))
ref = buf.narrow(0, 96, 32).sum(0)
print(f"node_sum 32 tracks x 16 renders: {t*1e3:.1f} us  {(33 * 16 * 2 * 131072 * 4) / t / 1e6:.0f} GB/s  max err {float((out[:, 0] - ref).abs().max()):.2e}")
t = timeit(lambda: F_.node_copy(x.transpose(0, 1), buf.narrow(0, 0, 32)))
print(f"node_copy 32 tracks x 16 renders: {t*1e3:.1f} us  {(2 * 32 * 16 * 2 * 131072 * 4) / t / 1e6:.0f} GB/s")
