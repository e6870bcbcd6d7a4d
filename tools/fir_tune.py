"""Times the FIR engine variants (CUDA events): long-filter partition size, mid-size FFT choice."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.functional as F_
from grafx_b200 import _cabi


def timeit(fn, warm=2, it=5):
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
B, L = 512, 131072
x = torch.randn(B, 2, L, device="cuda")
for Nh, key in ((96000, "long"), (60000, "long"), (4000, "mid"), (16384, "mid"), (1023, None), (2047, None), (400, None)):
    h = torch.randn(B, 2, Nh, device="cuda") / Nh ** 0.5
    opts = {"long": (4096, 8192, 16384), "mid": (8192, 16384), None: (0,)}[key]
    for n in opts:
        if key == "long":
            L_.gfx_fir_set_tuning(n, 0)
        elif key == "mid":
            L_.gfx_fir_set_tuning(0, n)
        ms = timeit(lambda: F_.fir_conv(x, h))
        print(f"taps {Nh:6d} n={L_.gfx_fir_fft_size(Nh):6d}: {ms:8.3f} ms  {B*2*L/ms/1e6:8.1f} Gsamples/s  {8*B*2*L/ms/1e6:7.0f} GB/s", flush=True)
L_.gfx_fir_set_tuning(8192, 8192)
h = torch.randn(B, 2, 96000, device="cuda") / 96000 ** 0.5
for mode, d in ((0, 0), (1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (1, 8)):
    L_.gfx_fir_set_long_mode(mode, d)
    ms = timeit(lambda: F_.fir_conv(x, h))
    print(f"long mode {mode} lookahead {d}: taps 96000: {ms:8.3f} ms", flush=True)
L_.gfx_fir_set_long_mode(0, 4)
