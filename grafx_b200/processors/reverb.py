"""STFTMaskedNoiseReverb -- drop-in for grafx.processors.reverb.STFTMaskedNoiseReverb
(reverb.py:15-228).  Same constructor kwargs / forward signature / parameter_size().

The fixed noise STFT is built once in __init__ exactly like upstream (numpy RandomState(0)
uniform noise -> torch.stft), so it is bit-identical to the reference's buffer.  Per call, the
impulse response is synthesised by csrc/reverb.cu and applied by the FIR engine (csrc/fir.cu)."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from .. import functional as F_


class STFTMaskedNoiseReverb(nn.Module):
    def __init__(self, ir_len=60000, processor_channel="pseudo_midside", n_fft=384, hop_length=192,
                 fixed_noise=True, gain_envelope=False, flashfftconv=True, max_input_len=2**17):
        super().__init__()
        if processor_channel not in ("mono", "stereo", "midside", "pseudo_midside"):
            raise ValueError(f"Invalid processor_channel: {processor_channel}")
        self.ir_len = ir_len
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.num_frames = 1 + (ir_len // hop_length)
        self.num_bins = 1 + n_fft // 2
        self.fixed_noise = fixed_noise
        self.gain_envelope = gain_envelope
        self.processor_channel = processor_channel
        self.register_buffer("window", torch.hann_window(n_fft))
        self.register_buffer("arange", torch.arange(self.num_frames).view(1, 1, 1, -1))  # (upstream buffer, reverb.py:77-78)
        if fixed_noise:
            rng = np.random.RandomState(0)
            noise = torch.tensor(rng.uniform(size=(2, ir_len)) * 2 - 1).float()
            self.register_buffer("noise_stft", self._stft(noise).unsqueeze(0))

    def _stft(self, noise):
        return torch.stft(noise, n_fft=self.n_fft, hop_length=self.hop_length, window=self.window.to(noise.device),
                          return_complex=True)

    def _noise(self, batch, device):
        if self.fixed_noise:
            return self.noise_stft
        # fresh noise per call (reverb.py:116-128): parameter-side preparation, O(batch * ir_len)
        noise = torch.rand(batch * 2, self.ir_len, device=device) * 2 - 1
        st = self._stft(noise)
        return st.reshape(batch, 2, st.shape[-2], st.shape[-1])

    def _synthesize(self, init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude, finish):
        genv = gain_env_log_magnitude if self.gain_envelope else None
        noise = self._noise(init_log_magnitude.shape[0], init_log_magnitude.device)
        return F_.reverb_ir(noise, init_log_magnitude, delta_log_magnitude, genv, self.window, self.ir_len,
                            self.n_fft, self.hop_length, finish=finish)

    def accepts_parameter_repeat(self):
        """render_grafx (4-D sources): this module takes the per-node parameter rows un-expanded (functional.shared_parameters)."""
        return bool(self.fixed_noise)

    def compute_ir(self, init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude=None):
        """The un-normalised mid/side impulse response [B, 2, ir_len], as upstream (reverb.py:161-200: the
        channel-mode epilogue and normalize_impulse belong to _process_*, not to compute_ir)."""
        return self._synthesize(init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude, "raw")[0]

    def forward(self, input_signals, init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude=None):
        # un-normalised response (+ row energies) in its final channel layout; normalize_impulse
        # (reverb.py:215-228) happens inside the convolution while the filter spectra are formed
        to_lr = self.processor_channel == "pseudo_midside"
        if F_._wants_grad(input_signals, init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude):
            return self._forward_training(input_signals, init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude)
        # render_grafx (4-D sources) repeats every node's parameters over the batch of renders: synthesise each
        # response (and its spectra) once -- the parameters arrive expanded ([N, ...]: every rep-th row is taken) or, when
        # render_grafx knows this module accepts them (`accepts_parameter_repeat`), as the un-expanded [N / rep, ...] rows
        rep = F_.parameter_repeat()
        if rep > 1 and self.fixed_noise and init_log_magnitude.shape[0] * rep == input_signals.shape[0]:
            pass
        elif rep > 1 and self.fixed_noise and init_log_magnitude.shape[0] % rep == 0:
            init_log_magnitude, delta_log_magnitude = init_log_magnitude[::rep], delta_log_magnitude[::rep]
            if gain_env_log_magnitude is not None:
                gain_env_log_magnitude = gain_env_log_magnitude[::rep]
        else:
            rep = 1
        ir_raw, energy = self._synthesize(init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude,
                                          "raw_lr" if to_lr else "raw")
        if self.processor_channel == "midside":
            return F_.ms_to_lr(F_.fir_conv_midside_ir(F_.lr_to_ms(input_signals), ir_raw, energy, to_lr=False, h_repeat=rep))
        return F_.fir_conv_midside_ir(input_signals, ir_raw, energy, to_lr=to_lr, h_repeat=rep)

    def _forward_training(self, input_signals, init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude):
        """Grad mode (grafx_b200/training.py): the response as upstream states it (PyTorch istft, O(parameters x frames)),
        the channel-mode epilogue of reverb.py:215-228, the convolution on the differentiable FIR engine."""
        from .. import training

        if F_._wants_grad(init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude):
            genv = gain_env_log_magnitude if self.gain_envelope else None
            noise = self._noise(init_log_magnitude.shape[0], init_log_magnitude.device)
            ir = training.stft_reverb_ir(noise, init_log_magnitude, delta_log_magnitude, genv, self.window, self.ir_len,
                                         self.n_fft, self.hop_length)
        else:  # only the audio needs a gradient: the synthesis kernel, no graph
            with torch.no_grad():
                ir = self.compute_ir(init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude)
        if self.processor_channel == "pseudo_midside":
            ir = torch.stack([ir[:, 0] + ir[:, 1], ir[:, 0] - ir[:, 1]], 1)
        ir = F_.normalize_impulse(ir)
        if self.processor_channel == "midside":
            return F_.ms_to_lr(F_.fir_conv(F_.lr_to_ms(input_signals), ir, "causal"))
        return F_.fir_conv(input_signals, ir, "causal")

    def parameter_size(self):
        size = {"init_log_magnitude": (2, self.num_bins), "delta_log_magnitude": (2, self.num_bins)}
        if self.gain_envelope:
            size["gain_env_log_magnitude"] = (2, self.num_frames)
        return size


def _filtered_noise(noise_len, num_channels, num_bands, f_min, f_max, scale, sr, zerophase, order):
    """Constructor-time host preprocessing, as upstream (core/noise.py:9-75): uniform noise from numpy's global
    generator split into `num_bands` bands by a Linkwitz-Riley tree of Butterworth sections (scipy)."""
    from scipy.signal import butter, sosfilt, sosfiltfilt

    from .core.fir import _from_scale, _to_scale

    x = 2 * np.random.rand(num_channels, noise_len) - 1
    s_breaks = np.linspace(_to_scale(f_min, scale), _to_scale(f_max, scale), num_bands * 2 - 1)[1::2]
    f_breaks = _from_scale(torch.from_numpy(s_breaks), scale).numpy()
    bands = []
    for f in f_breaks:
        lp, hp = butter(order, f, "lowpass", fs=sr, output="sos"), butter(order, f, "highpass", fs=sr, output="sos")
        if zerophase:
            low, x = sosfiltfilt(lp, x), sosfiltfilt(hp, x)
        else:
            low, x = sosfilt(lp, sosfilt(lp, x)), sosfilt(hp, sosfilt(hp, x))
        bands.append(low)
    bands.append(x)
    return torch.from_numpy(np.stack(bands, 1)).float()


class FilteredNoiseShapingReverb(nn.Module):
    """Drop-in for grafx.processors.reverb.FilteredNoiseShapingReverb (reverb.py:231-460): band-filtered noise shaped
    by per-band exponential decays (and optional fade-ins), unit-energy normalised, applied by causal convolution.
    The [B, C, K, T] envelope broadcast of the reference is one fused kernel (csrc/reverb.cu:
    noise_shaping_ir_kernel); the normalisation is folded into the filter spectra of the FIR engine."""

    def __init__(self, ir_len=60000, num_bands=12, processor_channel="midside", f_min=31.5, f_max=15000, scale="log",
                 sr=30000, zerophase=True, order=2, noise_randomness="pseudo-random", use_fade_in=False,
                 min_decay_ms=50, max_decay_ms=2000, flashfftconv=True, max_input_len=2**17):
        super().__init__()
        if processor_channel not in ("midside", "stereo", "mono"):
            raise ValueError(f"Unknown channel type: {processor_channel}")
        if noise_randomness not in ("pseudo-random", "fixed"):
            raise ValueError(f"Invalid noise_randomness argument: {noise_randomness}")
        self.num_bands = num_bands
        self.processor_channel = processor_channel
        self.num_channels = 1 if processor_channel == "mono" else 2
        self.ir_len = ir_len
        self.noise_randomness = noise_randomness
        noise_len = ir_len if noise_randomness == "fixed" else ir_len * 5
        self.register_buffer("filtered_noise", _filtered_noise(noise_len, self.num_channels, num_bands, f_min, f_max,
                                                               scale, sr, zerophase, order).unsqueeze(0))
        self.min_decay = (-60 / (min_decay_ms * sr / 1000)) / 20 * math.log(10)
        self.max_decay = (-60 / (max_decay_ms * sr / 1000)) / 20 * math.log(10)
        self.use_fade_in = use_fade_in
        self.register_buffer("arange", torch.arange(ir_len)[None, None, None, :])  # (upstream buffer, reverb.py:361-362)

    def compute_ir(self, log_decay, log_gain, log_fade_in=None, z_fade_in_gain=None):
        """Un-normalised response [B, C, ir_len] and its row energies [B, C]."""
        decay = torch.sigmoid(log_decay) * (self.max_decay - self.min_decay) + self.min_decay
        fade = fade_gain = None
        if self.use_fade_in:
            fade = torch.sigmoid(log_fade_in) * (decay - self.min_decay) + self.min_decay
            fade_gain = torch.sigmoid(z_fade_in_gain)
        start = 0
        if self.noise_randomness == "pseudo-random":
            start = int(torch.randint(0, self.filtered_noise.shape[-1] - self.ir_len, (1,)))
        return F_.noise_shaping_ir(self.filtered_noise[0], start, decay, log_gain, fade, fade_gain, self.ir_len)

    def forward(self, input_signals, log_decay, log_gain, log_fade_in=None, z_fade_in_gain=None):
        ir, energy = self.compute_ir(log_decay, log_gain, log_fade_in, z_fade_in_gain)
        if self.num_channels == 1:
            return F_.fir_conv(input_signals, ir * torch.rsqrt(energy + 1e-12).unsqueeze(-1), "causal")
        if self.processor_channel == "midside":
            return F_.ms_to_lr(F_.fir_conv_midside_ir(F_.lr_to_ms(input_signals), ir, energy, to_lr=False))
        return F_.fir_conv_midside_ir(input_signals, ir, energy, to_lr=False)

    def parameter_size(self):
        shape = (self.num_channels, self.num_bands)
        size = {"log_decay": shape, "log_gain": shape}
        if self.use_fade_in:
            size["log_fade_in"] = shape
            size["z_fade_in_gain"] = shape
        return size
