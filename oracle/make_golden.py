"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by EXECUTING THE REFERENCE's own code
(imported from /root/reference/src through oracle/ref_loader.py, even-pad guard on) in the
build container.  Run:  python -m oracle.make_golden        (needs /root/reference)

Each fixture holds: the input signal, the parameter tensors, the constructor kwargs (json) and
the reference output (float32).  Where the reference itself is broken at this commit the
fixture says so in `note` and the stored output comes from the nearest working reference path:
  - FIRFilter cannot be constructed: output = the reference's own
    normalize_impulse / tanh / FIRConvolution(causal) called in the order of filter.py:65-77;
  - IIRFilter(backend="ssm") is wrong for K>=2 (R2): the ssm fixture stores the lfilter output
    (the reference's known-answer test asserts the two are equal);
  - Ballistics depends on torchcomp (absent): produced with the shim recurrence in
    ref_loader.compressor_core_loop -- PARITY UNPINNED, flagged in `note`.
Also stores, for the FFT-convolution cases, the as-shipped (unguarded) output distance so the
R1 deviation stays visible.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")


def _params(size_dict, n, std, gen):
    out = {}
    for k, v in size_dict.items():
        shape = (v,) if isinstance(v, int) else tuple(v)
        out[k] = std * torch.randn(n, *shape, generator=gen)
    return out


def _save(name, x, params, kwargs, y, note="", extra=None):
    arrays = {"x": x.numpy(), "y": y.detach().float().numpy()}
    for k, v in params.items():
        arrays["p_" + k] = v.numpy()
    if extra:
        for k, v in extra.items():
            arrays["e_" + k] = np.asarray(v)
    meta = {"name": name, "kwargs": kwargs, "note": note}
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print(f"{name:40s} x{tuple(x.shape)} -> y{tuple(y.shape)}  {note}")


def main():
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle.ref_loader import load_reference

    load_reference(even_pad_guard=True)
    import grafx.processors as P
    from grafx.processors.core.convolution import FIRConvolution
    from grafx.processors.core.iir import IIRFilter
    from grafx.processors.core.utils import normalize_impulse
    from grafx.processors.core.midside import lr_to_ms, ms_to_lr

    os.makedirs(OUT, exist_ok=True)
    gen = torch.Generator().manual_seed(0)

    def randn(*s):
        return torch.randn(*s, generator=gen)

    with torch.no_grad():
        # ---------------- IIR family (exact + fsm)
        for backend in ("lfilter", "fsm"):
            for std in (0.01, 1.0):
                for ch in ("mono", "stereo", "midside"):
                    kw = dict(num_filters=5, processor_channel=ch, backend=backend, flashfftconv=False, fsm_fir_len=1000)
                    proc = P.ParametricEqualizer(**kw)
                    x = randn(2, 2, 2048)
                    prm = _params(proc.parameter_size(), 2, std, gen)
                    _save(f"peq_{backend}_{ch}_std{std}", x, prm, kw, proc(x, **prm))
        for cls in ("BiquadFilter", "StateVariableFilter"):  # PoleZeroFilter: shape bug upstream (Bs is 3-D)
            kw = dict(num_filters=3, backend="lfilter", flashfftconv=False)
            proc = getattr(P, cls)(**kw)
            x = randn(2, 2, 1500)
            prm = _params(proc.parameter_size(), 2, 0.3, gen)
            _save(f"{cls.lower()}_lfilter", x, prm, kw, proc(x, **prm))
        for cls in ("LowPassFilter", "HighPassFilter", "BandPassFilter", "BandRejectFilter", "AllPassFilter"):
            kw = dict(backend="lfilter", flashfftconv=False)
            proc = getattr(P, cls)(**kw)
            x = randn(2, 1, 1500)
            prm = _params(proc.parameter_size(), 2, 1.0, gen)
            _save(f"{cls.lower()}_lfilter", x, prm, kw, proc(x, **prm))
        for cls in ("PeakingFilter", "LowShelf", "HighShelf"):
            kw = dict(num_filters=2, backend="lfilter", flashfftconv=False)
            proc = getattr(P, cls)(**kw)
            x = randn(2, 2, 1500)
            prm = _params(proc.parameter_size(), 2, 1.0, gen)
            _save(f"{cls.lower()}_lfilter", x, prm, kw, proc(x, **prm))
        # config 1 of BASELINE.json at full size (1 x 2ch x 48000), both backends
        for backend in ("lfilter", "fsm"):
            kw = dict(num_filters=1, backend=backend, flashfftconv=False)
            proc = P.BiquadFilter(**kw)
            x = randn(1, 2, 48000)
            prm = _params(proc.parameter_size(), 1, 0.5, gen)
            _save(f"cfg1_biquad_{backend}", x, prm, kw, proc(x, **prm))
        # ssm fixture (R2: stores lfilter output), float64 KAT shape of test_filter.py:215-233 (reduced T)
        x = randn(7, 5, 1000).double()
        Bs = randn(7, 5, 6, 3).double()
        a1 = torch.rand(7, 5, 6, generator=gen).double() * 4 - 2
        a2 = ((torch.rand(7, 5, 6, generator=gen).double() * 2 - 1) * (2 - a1.abs()) + a1.abs()) * 0.5
        As = torch.stack([torch.ones_like(a1), a1, a2], -1)
        y = IIRFilter(backend="lfilter")(x, Bs, As)
        np.savez_compressed(os.path.join(OUT, "kat_iir_f64.npz"), x=x.numpy(), Bs=Bs.numpy(), As=As.numpy(), y=y.numpy())
        print("kat_iir_f64", tuple(y.shape))

        # ---------------- FIR / convolution
        conv = FIRConvolution(mode="causal", flashfftconv=False)
        for ch, nch in (("mono", 1), ("stereo", 2), ("midside", 2)):
            x = randn(2, 2, 2048)
            fir = 0.5 * randn(2, nch, 255)
            f = normalize_impulse(torch.tanh(fir))
            y = ms_to_lr(conv(lr_to_ms(x), f)) if ch == "midside" else conv(x, f)
            _save(f"firfilter_{ch}", x, {"fir": fir}, dict(fir_len=255, processor_channel=ch), y,
                  note="FIRFilter ctor broken upstream (R3): composed from reference functions per filter.py:65-77")
        from grafx.processors.core import convolution as cmod
        x = randn(2, 2, 3000)
        h = randn(2, 2, 300)
        y_g = cmod.convolve(x, h, mode="causal")
        y_z = cmod.convolve(x, h, mode="zerophase")
        cmod.compute_pad_len = cmod._compute_pad_len_as_shipped
        y_shipped = cmod.convolve(x, h, mode="causal")
        load_reference(even_pad_guard=True)
        _save("convolve_causal_zerophase", x, {"h": h}, {}, y_g, extra={"y_zerophase": y_z.numpy(),
              "as_shipped_rel_l2": float((y_shipped - y_g).norm() / y_g.norm())},
              note="even-pad guard; as_shipped_rel_l2 = distance of the unguarded reference (R1)")

        # ---------------- dynamics
        for cls in ("Compressor", "NoiseGate"):
            for es in ("iir", "ballistics", None):
                for gs, inlog in ((None, False), ("iir", False), ("iir", True), ("ballistics", False), ("ballistics", True)):
                    for knee in ("quadratic", "hard", "exponential"):
                        if gs is not None and knee != "quadratic":
                            continue
                        kw = dict(energy_smoother=es, gain_smoother=gs, gain_smooth_in_log=inlog, knee=knee,
                                  iir_len=256, flashfftconv=False)
                        proc = getattr(P, cls)(**kw)
                        x = randn(2, 2, 1024)
                        prm = _params(proc.parameter_size(), 2, 1.0, gen)
                        note = "PARITY UNPINNED (ballistics shim)" if "ballistics" in (es, gs) else ""
                        _save(f"{cls.lower()}_{es}_{gs}_{'log' if inlog else 'lin'}_{knee}", x, prm, kw, proc(x, **prm), note)
        # a slow one-pole (alpha^N not negligible) to exercise the truncation tail
        kw = dict(energy_smoother="iir", iir_len=256, flashfftconv=False)
        proc = P.Compressor(**kw)
        x = randn(2, 1, 3000)
        prm = _params(proc.parameter_size(), 2, 0.5, gen)
        prm["z_alpha_pre"] = torch.tensor([[6.0], [4.0]])
        _save("compressor_slow_pole", x, prm, kw, proc(x, **prm))

        # ---------------- reverb
        for ch in ("pseudo_midside", "midside", "stereo"):
            for genv in (False, True):
                kw = dict(ir_len=1920, processor_channel=ch, gain_envelope=genv, flashfftconv=False)
                proc = P.STFTMaskedNoiseReverb(**kw)
                x = randn(2, 2, 2048)
                prm = _params(proc.parameter_size(), 2, 0.5, gen)
                ir = proc.compute_ir(**prm)
                _save(f"reverb_{ch}_{'genv' if genv else 'plain'}", x, prm, kw, proc(x, **prm), extra={"ir": ir.numpy()})

        # ---------------- graph render (tests/graph/test_render.py:13-37 shape, real processors)
        from grafx.data import GRAFX, NodeConfigs, convert_to_tensor
        from grafx.render import prepare_render, render_grafx, reorder_for_fast_render

        config = NodeConfigs(["eq", "compressor", "reverb"])
        G = GRAFX(config=config)
        out_id = G.add("out")
        for _ in range(3):
            _, end_id = G.add_serial_chain(["in", "eq", "compressor", "reverb"])
            G.connect(end_id, out_id)
        G_t = reorder_for_fast_render(convert_to_tensor(G), method="beam")
        rd = prepare_render(G_t)
        kws = {"eq": dict(num_filters=5, processor_channel="stereo", backend="lfilter", flashfftconv=False),
               "compressor": dict(flashfftconv=False, iir_len=256),
               "reverb": dict(ir_len=1920, flashfftconv=False)}
        procs = {"eq": P.ParametricEqualizer(**kws["eq"]), "compressor": P.Compressor(**kws["compressor"]),
                 "reverb": P.STFTMaskedNoiseReverb(**kws["reverb"])}
        params = {t: _params(procs[t].parameter_size(), 3, 0.3, gen) for t in procs}
        plan = {"num_nodes": int(rd.num_nodes), "max_order": int(rd.max_order), "iters": []}
        for it in rd.iter_list:
            def acc(a):
                idx = a.idx
                if isinstance(idx, torch.Tensor):
                    idx = idx.tolist()
                return [a.method, list(idx)]
            plan["iters"].append({"type": it.node_type, "reads": [acc(a) for a in it.source_reads],
                                  "aggs": [[a.method, (a.idx.tolist() if a.idx is not None else None)] for a in it.aggregations],
                                  "param": acc(it.parameter_read), "write": acc(it.dest_write)})
        for tag, xin in (("3d", randn(3, 2, 1024)), ("4d", randn(2, 3, 2, 1024))):
            out, _, buf = render_grafx(procs, xin, params, rd, parameters_grad=False)
            arrays = {"x": xin.numpy(), "y": out.numpy(), "buffer": buf.numpy(),
                      "meta": np.frombuffer(json.dumps({"plan": plan, "kwargs": kws}).encode(), dtype=np.uint8)}
            for t in params:
                for k, v in params[t].items():
                    arrays[f"p_{t}__{k}"] = v.numpy()
            np.savez_compressed(os.path.join(OUT, f"render_mix3_{tag}.npz"), **arrays)
            print("render_mix3_" + tag, tuple(out.shape), tuple(buf.shape))


if __name__ == "__main__":
    main()
