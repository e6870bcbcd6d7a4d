"""Helpers shared by the oracle (CPU) and parity (GPU) tests: fixture loading and the mapping
from a fixture name to (a) the oracle call and (b) the grafx_b200 processor."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CLASS_OF_PREFIX = {
    "peq": "ParametricEqualizer",
    "cfg1_biquad": "BiquadFilter",
    "biquadfilter": "BiquadFilter",
    "statevariablefilter": "StateVariableFilter",
    "lowpassfilter": "LowPassFilter",
    "highpassfilter": "HighPassFilter",
    "bandpassfilter": "BandPassFilter",
    "bandrejectfilter": "BandRejectFilter",
    "allpassfilter": "AllPassFilter",
    "peakingfilter": "PeakingFilter",
    "lowshelf": "LowShelf",
    "highshelf": "HighShelf",
    "firfilter": "FIRFilter",
    "compressor": "Compressor",
    "noisegate": "NoiseGate",
    "reverb": "STFTMaskedNoiseReverb",
    # SURVEY.md section 8(f) "next" rows (oracle/make_golden_next.py)
    "next_geq": "GraphicEqualizer",
    "next_zpfir_old": "ZeroPhaseFIREqualizer",
    "next_zpfir": "NewZeroPhaseFIREqualizer",
    "next_stereogain": "StereoGain",
    "next_sidegainimager": "SideGainImager",
    "next_tanhdistortion": "TanhDistortion",
    "next_piecewisetanhdistortion": "PiecewiseTanhDistortion",
    "next_powerdistortion": "PowerDistortion",
    "next_chebyshevdistortion": "ChebyshevDistortion",
    "next_parallelmix": "ParallelMix",
    "next_multitapdelay": "MultitapDelay",
    "next_noiseshapingreverb": "FilteredNoiseShapingReverb",
    "next_approxcompressor": "ApproxCompressor",
    "next_approxnoisegate": "ApproxNoiseGate",
    "next_gainstaging": "GainStagingRegularization",
    "next_stftreverb": "STFTMaskedNoiseReverb",
}


def fixture_names(prefixes=None):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    names = [n for n in names if not n.startswith(("kat_", "render_", "convolve_", "envelope_", "fullsize_"))]
    if not prefixes or not any(p.startswith("grad_") for p in prefixes):
        names = [n for n in names if not n.startswith("grad_")]  # gradient fixtures (oracle/make_golden_grad.py)
    if prefixes:
        names = [n for n in names if n.startswith(tuple(prefixes))]
    return names


def class_of(name):
    for pre in sorted(CLASS_OF_PREFIX, key=len, reverse=True):
        if name.startswith(pre):
            return CLASS_OF_PREFIX[pre]
    raise KeyError(name)


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode()) if "meta" in z.files else {}
    x = torch.from_numpy(z["x"])
    y = torch.from_numpy(z["y"])
    params = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p_")}
    extra = {k[2:]: z[k] for k in z.files if k.startswith("e_")}
    return x, params, meta, y, extra


def oracle_call(name, x, params, kwargs, dtype=None, extra=None):
    """Runs the oracle restatement for the fixture `name` (`extra`: the constants the reference registered)."""
    from oracle import grafx_oracle as O

    if dtype is not None:
        x = x.to(dtype)
        params = {k: v.to(dtype) for k, v in params.items()}
    cls = class_of(name)
    kw = dict(kwargs)
    ex = {k: torch.from_numpy(np.asarray(v)) for k, v in (extra or {}).items()}
    if cls == "GraphicEqualizer":
        return O.graphic_equalizer(x, params["log_gains"], ex["fc"], ex["fB"], sr=kw["sr"],
                                   processor_channel=kw["processor_channel"], use_torchaudio=dtype is None)
    if cls in ("ZeroPhaseFIREqualizer", "NewZeroPhaseFIREqualizer"):
        return O.zerophase_fir_equalizer(x, params["log_magnitude"], ex.get("window"), ex.get("filterbank"),
                                         processor_channel=kw.get("processor_channel", "mono"))
    if cls == "StereoGain":
        return O.stereo_gain(x, params["log_gain"])
    if cls == "SideGainImager":
        return O.side_gain_imager(x, params["log_gain"])
    if cls == "TanhDistortion":
        return O.tanh_distortion(x, params.get("log_pre_gain"), params.get("log_post_gain"), params.get("bias"),
                                 inverse_post_gain=kw.get("inverse_post_gain", True), remove_dc=kw.get("remove_dc", False))
    if cls == "PiecewiseTanhDistortion":
        return O.piecewise_tanh_distortion(x, params["log_hardness"], params["z_threshold"], params.get("log_pre_gain"),
                                           params.get("log_post_gain"), inverse_post_gain=kw.get("inverse_post_gain", True),
                                           remove_dc=kw.get("remove_dc", False))
    if cls in ("PowerDistortion", "ChebyshevDistortion"):
        return O.series_distortion("power" if cls == "PowerDistortion" else "chebyshev", x, params["basis_weights"],
                                   params.get("log_pre_gain"), remove_dc=kw.get("remove_dc", False),
                                   use_tanh=kw.get("use_tanh", False))
    if cls == "MultitapDelay":
        nch = 1 if kw["processor_channel"] == "mono" else 2
        return O.multitap_delay(x, params["delay_z"], params.get("log_fir_magnitude"), ex.get("window"),
                                segment_len=kw["segment_len"], num_segments=kw["num_segments"],
                                num_delay_per_segment=kw["num_delay_per_segment"], num_channels=nch,
                                pre_delay=kw.get("pre_delay", 0))[0]
    if cls == "FilteredNoiseShapingReverb":
        return O.noise_shaping_reverb(x, params["log_decay"], params["log_gain"], ex["filtered_noise"], float(ex["min_decay"]),
                                      float(ex["max_decay"]), params.get("log_fade_in"), params.get("z_fade_in_gain"),
                                      processor_channel=kw["processor_channel"])
    if cls in ("ApproxCompressor", "ApproxNoiseGate"):
        fn = O.approx_compressor if cls == "ApproxCompressor" else O.approx_noisegate
        return fn(x, params["z_alpha"], params["log_threshold"], params["log_ratio"], params["log_knee"],
                  iir_len=kw["iir_len"])
    if cls == "GainStagingRegularization":
        return O.stereo_gain(x, params["log_gain"])
    if cls == "ParallelMix":
        w = O.parallel_mix_weights(params["parallel_weights"], kw["activation"])
        a = O.tanh_distortion(x, params["a__log_pre_gain"])
        b = O.stereo_gain(x, params["b__log_gain"])
        return a * w[:, 0, None, None] + b * w[:, 1, None, None]
    backend = kw.get("backend", "lfilter")
    fir_len = kw.get("fsm_fir_len", 4000)
    ta = dtype is None

    def run_iir(Bs, As):
        return O._iir(x, Bs, As, backend, fir_len, use_torchaudio=ta)

    if cls == "ParametricEqualizer":
        return O.parametric_equalizer(x, **params, processor_channel=kw["processor_channel"], backend=backend,
                                      fsm_fir_len=fir_len, use_torchaudio=ta)
    if cls == "BiquadFilter":
        return O.biquad_filter(x, **params, backend=backend, fsm_fir_len=fir_len, use_torchaudio=ta)
    if cls == "StateVariableFilter":
        Bs, As = O.svf_coeffs(**params)
        return run_iir(Bs.unsqueeze(1), As.unsqueeze(1))
    if cls in ("LowPassFilter", "HighPassFilter", "BandPassFilter", "BandRejectFilter", "AllPassFilter"):
        c, alpha, _ = O.peq_activations(params["w0"], params["q_inv"])
        Bs, As = O.simple_filter_coeffs(cls[:-6].lower(), c, alpha)
        return run_iir(Bs.unsqueeze(1), As.unsqueeze(1))
    if cls in ("PeakingFilter", "LowShelf", "HighShelf"):
        c, alpha, A = O.peq_activations(params["w0"], params["q_inv"], params["log_gain"])
        fn = {"PeakingFilter": O.peaking_coeffs, "LowShelf": O.lowshelf_coeffs, "HighShelf": O.highshelf_coeffs}[cls]
        Bs, As = fn(c, alpha, A)
        return run_iir(Bs.unsqueeze(1), As.unsqueeze(1))
    if cls == "FIRFilter":
        return O.fir_filter(x, params["fir"], kw["processor_channel"])
    if cls in ("Compressor", "NoiseGate"):
        kw.pop("flashfftconv", None)
        return O.dynamics(cls.lower(), x, **params, **kw)
    if cls == "STFTMaskedNoiseReverb":
        return O.stft_masked_noise_reverb(x, **params, ir_len=kw["ir_len"], processor_channel=kw["processor_channel"],
                                          n_fft=kw.get("n_fft", 384), hop=kw.get("hop_length", 192))
    raise KeyError(cls)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def oracle_gradients(name, x, params, kw, w):
    """float64 gradients of sum(w * y) through the oracle restatement (torchaudio lfilter / torch.fft autograd) for
    the gradient fixture `name` ("grad_<forward fixture name>").  Returns (y, gx, {param: grad})."""
    x64 = x.double().requires_grad_(True)
    p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
    y = oracle_call(name[len("grad_"):], x64, p64, kw)
    (y * w.double()).sum().backward()
    return y.detach(), x64.grad, {k: v.grad for k, v in p64.items()}
