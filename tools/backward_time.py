"""Device time of one training step (forward + backward of sum(w * y)) of ParametricEqualizer at the config-2 size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.processors as P

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
backend = sys.argv[2] if len(sys.argv) > 2 else "lfilter"
torch.manual_seed(0)
proc = P.ParametricEqualizer(num_filters=5, processor_channel="stereo", backend=backend).cuda()
x = torch.randn(B, 2, 131072, device="cuda")
w = torch.randn_like(x)
prm = {k: (0.3 * torch.randn(B, *v, device="cuda")).requires_grad_(True) for k, v in proc.parameter_size().items()}

def step():
    for p in prm.values():
        p.grad = None
    y = proc(x, **prm)
    (y * w).sum().backward()

def fwd():
    with torch.no_grad():
        proc(x, **prm)

for name, fn in (("forward (no grad)", fwd), ("forward + backward", step)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): fn()
    b.record(); torch.cuda.synchronize()
    print(f"PEQ K=5 stereo {backend} B={B} x 2 x 131072: {name} {a.elapsed_time(b) / 5:.3f} ms")
