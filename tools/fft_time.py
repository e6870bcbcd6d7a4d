"""Event timings of the FFT-engine workloads: FIR 1023 (single-partition overlap-save), reverb-shape long filter,
fsm-length filter.  GRAFX_B200_LIB selects a library variant."""
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


tag = os.path.basename(os.environ.get("GRAFX_B200_LIB", "default"))
x = torch.randn(512, 2, 131072, device="cuda")
res = []
for N in (1023, 4000, 400, 96000):
    h = torch.randn(512, 2, N, device="cuda") / N ** 0.5
    res.append(f"{N} taps {timeit(lambda: F_.fir_conv(x, h)):.4f} ms")
print(f"{tag:20s} " + " | ".join(res), flush=True)
