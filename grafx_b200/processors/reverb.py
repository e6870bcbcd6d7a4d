"""STFTMaskedNoiseReverb -- drop-in for grafx.processors.reverb.STFTMaskedNoiseReverb
(reverb.py:15-228).  Same constructor kwargs / forward signature / parameter_size().

The fixed noise STFT is built once in __init__ exactly like upstream (numpy RandomState(0)
uniform noise -> torch.stft), so it is bit-identical to the reference's buffer.  Per call, the
impulse response is synthesised by csrc/reverb.cu and applied by the FIR engine (csrc/fir.cu)."""
from __future__ import annotations

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

    def compute_ir(self, init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude=None, _finish="unit"):
        """Mid/side impulse response.  NB upstream returns it un-normalised; here the kernel also
        applies the channel-mode epilogue, selected by `_finish`: "unit" keeps upstream's meaning up
        to the unit-energy scale, "lr" adds ms_to_lr, "raw" returns (raw response, row energies)."""
        genv = gain_env_log_magnitude if self.gain_envelope else None
        noise = self._noise(init_log_magnitude.shape[0], init_log_magnitude.device)
        return F_.reverb_ir(noise, init_log_magnitude, delta_log_magnitude, genv, self.window, self.ir_len,
                            self.n_fft, self.hop_length, finish=_finish)

    def forward(self, input_signals, init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude=None):
        # un-normalised response (+ row energies) in its final channel layout; normalize_impulse
        # (reverb.py:215-228) happens inside the convolution while the filter spectra are formed
        to_lr = self.processor_channel == "pseudo_midside"
        # render_grafx (4-D sources) repeats every node's parameters over the batch of renders: synthesise each
        # response (and its spectra) once
        rep = F_.parameter_repeat()
        if rep > 1 and self.fixed_noise and init_log_magnitude.shape[0] % rep == 0:
            init_log_magnitude, delta_log_magnitude = init_log_magnitude[::rep], delta_log_magnitude[::rep]
            if gain_env_log_magnitude is not None:
                gain_env_log_magnitude = gain_env_log_magnitude[::rep]
        else:
            rep = 1
        ir_raw, energy = self.compute_ir(init_log_magnitude, delta_log_magnitude, gain_env_log_magnitude,
                                         _finish="raw_lr" if to_lr else "raw")
        if self.processor_channel == "midside":
            return F_.ms_to_lr(F_.fir_conv_midside_ir(F_.lr_to_ms(input_signals), ir_raw, energy, to_lr=False, h_repeat=rep))
        return F_.fir_conv_midside_ir(input_signals, ir_raw, energy, to_lr=to_lr, h_repeat=rep)

    def parameter_size(self):
        size = {"init_log_magnitude": (2, self.num_bins), "delta_log_magnitude": (2, self.num_bins)}
        if self.gain_envelope:
            size["gain_env_log_magnitude"] = (2, self.num_frames)
        return size
