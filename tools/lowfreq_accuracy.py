"""fp32 cascade kernel vs torchaudio fp32 lfilter vs float64 truth (same fp32 coefficients) for realistic
low-frequency EQ bands at 48 kHz."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.functional as F_
from oracle import grafx_oracle as O

torch.manual_seed(0)
L = 131072
x = torch.randn(4, 1, L)
sr = 48000.0
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
print("kind    f0[Hz]    Q   | ref32-vs-truth  kernel32-vs-truth  kernel32-vs-ref32")
for kind in ("peak", "lowshelf"):
    for f0 in (20, 50, 100, 200, 500, 1000, 4000):
        for Q in (0.7, 4.0):
            w0 = 2 * math.pi * f0 / sr
            A = 10 ** (6 / 40)
            alpha = math.sin(w0) / (2 * Q)
            c = math.cos(w0)
            if kind == "peak":
                b = [1 + alpha * A, -2 * c, 1 - alpha * A]; a = [1 + alpha / A, -2 * c, 1 - alpha / A]
            else:
                S = 2 * math.sqrt(A) * alpha
                b = [A * ((A + 1) - (A - 1) * c + S), 2 * A * ((A - 1) - (A + 1) * c), A * ((A + 1) - (A - 1) * c - S)]
                a = [(A + 1) + (A - 1) * c + S, -2 * ((A - 1) + (A + 1) * c), (A + 1) + (A - 1) * c - S]
            Bs = torch.tensor(b, dtype=torch.float32).view(1, 1, 1, 3).expand(4, 1, 1, 3).contiguous()
            As = torch.tensor(a, dtype=torch.float32).view(1, 1, 1, 3).expand(4, 1, 1, 3).contiguous()
            y_ref = O.iir_lfilter(x, Bs, As, use_torchaudio=True)
            # truth: the filter torchaudio actually runs = fp32-normalised coefficients, evaluated in float64
            nb = (Bs / As[..., :1]).double(); na = (As / As[..., :1]).double()
            y64 = O.iir_lfilter(x.double(), nb, na, use_torchaudio=False)
            y = F_.biquad_cascade(x.cuda(), Bs.cuda(), As.cuda()).cpu()
            print(f"{kind:8s} {f0:6d} {Q:5.1f} |   {rel(y_ref, y64):.2e}        {rel(y, y64):.2e}           {rel(y, y_ref):.2e}")
