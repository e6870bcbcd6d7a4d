"""TEST INFRASTRUCTURE -- golden vectors of the stand-alone envelope modules, produced by EXECUTING THE REFERENCE's
own TruncatedOnePoleIIRFilter / Ballistics (core/envelope.py:10-101) and IIREnvelopeFollower /
BallisticsEnvelopeFollower (dynamics.py:745-790) through oracle/ref_loader.py (even-pad guard on; torchcomp replaced by
the stand-in documented in oracle/torchcomp_core.py).  Run:  python -m oracle.make_golden_envelope
Writes tests/golden/envelope_modules.npz."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle.ref_loader import load_reference

    load_reference(even_pad_guard=True)
    from grafx.processors.core.envelope import Ballistics, TruncatedOnePoleIIRFilter
    from grafx.processors.dynamics import BallisticsEnvelopeFollower, IIREnvelopeFollower

    g = torch.Generator().manual_seed(11)
    B, C, L, iir_len = 4, 2, 6000, 2048
    u = torch.rand(B, L, generator=g) * 2
    x = torch.randn(B, C, L, generator=g)
    z1 = torch.tensor([[4.0], [0.2], [8.5], [-1.0]])
    z2 = torch.randn(B, 2, generator=g)
    out = {"u": u.numpy(), "x": x.numpy(), "z1": z1.numpy(), "z2": z2.numpy(), "iir_len": np.int64(iir_len)}
    with torch.no_grad():
        out["y_onepole"] = TruncatedOnePoleIIRFilter(iir_len=iir_len, flashfftconv=False)(u, z1).numpy()
        out["y_ballistics"] = Ballistics()(u, z2).numpy()
        for det in ("energy", "amplitude"):
            out[f"env_iir_{det}"] = IIREnvelopeFollower(detect_with=det, iir_len=iir_len, flashfftconv=False)(x, z1).numpy()
            out[f"env_ballistics_{det}"] = BallisticsEnvelopeFollower(detect_with=det)(x, z2).numpy()
    path = os.path.join(HERE, "..", "tests", "golden", "envelope_modules.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
