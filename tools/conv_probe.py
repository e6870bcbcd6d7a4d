"""One long-filter convolution at the BASELINE reverb shape for a given partition size (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.functional as F_
from grafx_b200 import _cabi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
L_ = _cabi.lib()
L_.gfx_fir_set_tuning(n, 0)
x = torch.randn(B, 2, 131072, device="cuda")
h = torch.randn(B, 2, 96000, device="cuda") / 300
for _ in range(2):
    y = F_.fir_conv(x, h)
torch.cuda.synchronize()
print("ok", n, float(y.abs().mean()))
