import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.functional as F_
from grafx_b200.processors import design
torch.manual_seed(0)
B, C, L, K = 256, 2, 131072, int(sys.argv[1]) if len(sys.argv) > 1 else 5
x = torch.randn(B, C, L, device="cuda")
w0, q, g = (torch.randn(B, C, K, device="cuda") for _ in range(3))
Bs, As = design.parametric_eq(w0, q, g, use_shelving_filters=K >= 3)
for _ in range(4):
    y = F_.biquad_cascade(x, Bs, As)
torch.cuda.synchronize()
