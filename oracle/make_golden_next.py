"""TEST INFRASTRUCTURE -- golden vectors for the SURVEY.md section 8(f) "next" processors, produced by EXECUTING
THE REFERENCE's own code (oracle/ref_loader.py, even-pad guard on).  Run:  python -m oracle.make_golden_next
Writes tests/golden/next_*.npz (same layout as oracle/make_golden.py: x, p_<param>, e_<extra>, y, meta).
Constants the reference registers as buffers (band tables, windows, filterbank matrices) are stored as extras so
that the oracle restatement is pinned against them instead of carrying its own copy."""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle.make_golden import _params, _save
    from oracle.ref_loader import load_reference

    load_reference(even_pad_guard=True)
    import grafx.processors as P

    gen = torch.Generator().manual_seed(1)

    def randn(*s):
        return torch.randn(*s, generator=gen)

    with torch.no_grad():
        # ---- GraphicEqualizer: K = 24 / 31 peaking sections in one cascade
        for scale in ("bark", "third_octave"):
            for ch in ("mono", "stereo", "midside"):
                kw = dict(processor_channel=ch, scale=scale, sr=44100, backend="lfilter", flashfftconv=False)
                proc = P.GraphicEqualizer(**kw)
                x = randn(2, 2, 2048)
                prm = _params(proc.parameter_size(), 2, 0.5, gen)
                _save(f"next_geq_{scale}_{ch}", x, prm, kw, proc(x, **prm),
                      extra={"fc": proc.geq.fc.numpy(), "fB": proc.geq.fB.numpy()})
        # ---- zero-phase FIR equalizers
        kw = dict(num_magnitude_bins=64)
        proc = P.ZeroPhaseFIREqualizer(**kw)
        x = randn(2, 2, 2048)
        prm = _params(proc.parameter_size(), 2, 0.5, gen)
        _save("next_zpfir_old", x, prm, kw, proc(x, **prm), extra={"window": proc.fir.window.numpy()})
        for ch in ("mono", "stereo", "midside"):
            for fb in (False, True):
                kw = dict(num_frequency_bins=64, processor_channel=ch, use_filterbank=fb,
                          filterbank_kwargs=dict(num_filters=20, f_max=22050) if fb else {}, flashfftconv=False)
                proc = P.NewZeroPhaseFIREqualizer(**kw)
                x = randn(2, 2, 2048)
                prm = _params(proc.parameter_size(), 2, 0.5, gen)
                extra = {"window": proc.fir.window.numpy()}
                if fb:
                    extra["filterbank"] = proc.fir.filterbank.filterbank.numpy()
                _save(f"next_zpfir_{ch}_{'fb' if fb else 'plain'}", x, prm, kw, proc(x, **prm), extra=extra)
        # ---- stereo utilities
        for cls in ("StereoGain", "SideGainImager"):
            proc = getattr(P, cls)()
            x = randn(3, 2, 1500)
            prm = _params(proc.parameter_size(), 3, 0.5, gen)
            _save(f"next_{cls.lower()}", x, prm, {}, proc(x, **prm))
        # ---- memoryless distortions
        for i, kw in enumerate((dict(), dict(inverse_post_gain=False, use_bias=True, remove_dc=True), dict(pre_post_gain=False))):
            proc = P.TanhDistortion(**kw)
            x = randn(3, 2, 1500)
            prm = _params(proc.parameter_size(), 3, 0.7, gen)
            _save(f"next_tanhdistortion_{i}", x, prm, kw, proc(x, **prm))
        for i, kw in enumerate((dict(), dict(inverse_post_gain=False, remove_dc=True))):
            proc = P.PiecewiseTanhDistortion(**kw)
            x = 2.0 * randn(3, 2, 1500)
            prm = _params(proc.parameter_size(), 3, 0.7, gen)
            _save(f"next_piecewisetanhdistortion_{i}", x, prm, kw, proc(x, **prm))
        for cls in ("PowerDistortion", "ChebyshevDistortion"):
            for i, kw in enumerate((dict(max_order=6), dict(max_order=10, pre_gain=False, remove_dc=True, use_tanh=True))):
                proc = getattr(P, cls)(**kw)
                x = 0.5 * randn(3, 2, 1500)
                prm = _params(proc.parameter_size(), 3, 0.5, gen)
                _save(f"next_{cls.lower()}_{i}", x, prm, kw, proc(x, **prm))
        # ---- MultitapDelay (surrogate delay lines + zero-phase colouring per tap)
        for i, kw in enumerate((dict(segment_len=300, num_segments=4, num_delay_per_segment=1, processor_channel="stereo", zp_filter_bins=8),
                                dict(segment_len=256, num_segments=3, num_delay_per_segment=2, processor_channel="mono",
                                     zp_filter_per_tap=False, pre_delay=17))):
            proc = P.MultitapDelay(**kw, flashfftconv=False)
            x = randn(2, 2 if kw["processor_channel"] == "stereo" else 1, 3000)
            prm = _params(proc.parameter_size(), 2, 1.0, gen)
            y, _ = proc(x, **prm)
            extra = {"window": proc.zp_filter.window.numpy()} if kw.get("zp_filter_per_tap", True) else {}
            _save(f"next_multitapdelay_{i}", x, prm, kw, y, extra=extra)
        # ---- FilteredNoiseShapingReverb (the noise buffer comes from numpy's global generator: seed stored)
        import numpy as np
        for i, kw in enumerate((dict(ir_len=2000, num_bands=4, processor_channel="stereo", noise_randomness="fixed"),
                                dict(ir_len=1500, num_bands=6, processor_channel="midside", noise_randomness="fixed", use_fade_in=True),
                                dict(ir_len=1800, num_bands=5, processor_channel="mono", noise_randomness="fixed", zerophase=False,
                                     scale="bark_traunmuller", f_max=12000))):
            np.random.seed(100 + i)
            proc = P.FilteredNoiseShapingReverb(**kw, flashfftconv=False)
            x = randn(2, 1 if kw["processor_channel"] == "mono" else 2, 3000)
            prm = _params(proc.parameter_size(), 2, 1.0, gen)
            _save(f"next_noiseshapingreverb_{i}", x, prm, dict(kw, numpy_seed=100 + i), proc(x, **prm),
                  extra={"filtered_noise": proc.filtered_noise[0].numpy(), "min_decay": proc.min_decay, "max_decay": proc.max_decay})
        # ---- ParallelMix of two distortions
        for act in ("softmax", "softplus"):
            proc = P.ParallelMix({"a": P.TanhDistortion(), "b": P.StereoGain()}, activation=act)
            x = randn(3, 2, 1500)
            size = proc.parameter_size()
            prm = {"parallel_weights": 0.5 * randn(3, size["parallel_weights"])}
            sub = {k: _params(size[k], 3, 0.5, gen) for k in ("a", "b")}
            y, _ = proc(x, prm["parallel_weights"], **sub)
            flat = dict(prm)
            for k, d in sub.items():
                for n, v in d.items():
                    flat[f"{k}__{n}"] = v
            _save(f"next_parallelmix_{act}", x, flat, dict(activation=act), y)

        # ---- ApproxCompressor / ApproxNoiseGate (deprecated upstream, still exported) and the gain-staging term
        for cls, key in (("ApproxCompressor", "iir_len"), ("ApproxNoiseGate", "freq_sample_n")):
            kw = {key: 1024, "flashfftconv": False}
            proc = getattr(P, cls)(**kw)
            x = randn(3, 2, 3000) * torch.tensor([0.02, 0.3, 1.0])[:, None, None]
            prm = _params(proc.parameter_size(), 3, 1.0, gen)
            prm["log_threshold"] = prm["log_threshold"] - 2.0  # put the knee inside the range of the envelopes
            _save(f"next_{cls.lower()}", x, prm, dict(iir_len=1024), proc(x, **prm))
        proc = P.GainStagingRegularization(P.StereoGain())
        x = randn(3, 2, 1500)
        prm = _params(proc.parameter_size(), 3, 0.5, gen)
        y, inter = proc(x, **prm)
        _save("next_gainstaging", x, prm, {}, y, extra={"gain_reg": inter["gain_reg"].numpy()})

        # ---- STFTMaskedNoiseReverb with non-default STFT geometries (n_fft / hop_length constructor arguments)
        for n_fft, hop, ch, genv in ((256, 64, "pseudo_midside", False), (512, 256, "midside", True), (128, 96, "stereo", False),
                                     (1024, 256, "pseudo_midside", True)):
            kw = dict(ir_len=3000, processor_channel=ch, n_fft=n_fft, hop_length=hop, gain_envelope=genv, flashfftconv=False)
            proc = P.STFTMaskedNoiseReverb(**kw)
            x = randn(2, 2, 2048)
            prm = _params(proc.parameter_size(), 2, 0.5, gen)
            _save(f"next_stftreverb_{n_fft}_{hop}", x, prm, kw, proc(x, **prm))


if __name__ == "__main__":
    main()
