"""Error of the cascade kernel (fp32 / fp64) and of the reference (fp32 torchaudio lfilter) against the float64
oracle on the GraphicEqualizer fixtures (ill-conditioned low-frequency sections)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from _golden import load, oracle_call, rel_l2, fixture_names
import grafx_b200.processors as P
import grafx_b200.functional as F_
from oracle import grafx_oracle as O

for name in fixture_names(["next_geq"]):
    x, params, meta, y_ref, extra = load(name)
    kw = meta["kwargs"]
    y64 = oracle_call(name, x, params, kw, dtype=torch.float64, extra=extra)
    proc = P.GraphicEqualizer(**kw).cuda()
    y = proc(x.cuda(), **{k: v.cuda() for k, v in params.items()}).cpu()
    Bs, As = proc.geq(params["log_gains"].cuda())
    xin = O.lr_to_ms(x) if kw["processor_channel"] == "midside" else x
    yd = F_.biquad_cascade(xin.double().cuda(), Bs.double(), As.double()).cpu()
    if kw["processor_channel"] == "midside":
        yd = O.ms_to_lr(yd)
    print(f"{name:32s} ref32-vs-truth {rel_l2(y_ref, y64):.2e}   kernel32-vs-truth {rel_l2(y, y64):.2e}   kernel32-vs-ref {rel_l2(y, y_ref):.2e}   kernel64-vs-truth {rel_l2(yd, y64):.2e}")
