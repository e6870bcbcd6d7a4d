"""Event timings of the long-filter convolution: BASELINE reverb shape (per-item filters) and the config-5 shape
(32 filters shared by 16 renders each).  GRAFX_B200_LIB selects a library variant; argv: mac forms / partition sizes."""
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
tag = os.path.basename(os.environ.get("GRAFX_B200_LIB", "default"))
x = torch.randn(512, 2, 131072, device="cuda")
h = torch.randn(512, 2, 96000, device="cuda") / 300
h32 = torch.randn(32, 2, 96000, device="cuda") / 300
h60 = torch.randn(512, 2, 60000, device="cuda") / 300
combos = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(4096, 2), (4096, 1), (8192, 2), (8192, 1)]
for n, form in combos:
    L_.gfx_fir_set_tuning(n, 0)
    L_.gfx_fir_set_mac_form(form)
    a = timeit(lambda: F_.fir_conv(x, h))
    b = timeit(lambda: F_.fir_conv(x, h32, h_repeat=16))
    c = timeit(lambda: F_.fir_conv(x, h60))
    print(f"{tag:18s} n={n} mac_form={form}: 96000 taps {a:7.3f} ms | shared filters (32 x 16) {b:7.3f} ms | 60000 taps {c:7.3f} ms", flush=True)
