"""The stock-library GPU path for configs 2 and 3b on the same B200, for context: what the reference's own code runs
when its tensors live on a GPU -- K x torchaudio.functional.lfilter (CUDA: one thread per row, sequential in time,
core/iir.py:154-184) and the torch.fft convolution of core/convolution.py:119-134.  Library calls only; none of this
repository's kernels.  usage: python tools/stock_gpu_baseline.py"""
import time
import torch
import torchaudio.functional as AF


def timed(fn, warm=1, it=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


torch.manual_seed(0)
# config 2: 256 x 2 x 131072, K = 5 sections (stable random biquads; the cost does not depend on the values)
B, C, L, K = 256, 2, 131072, 5
x = torch.randn(B, C, L, device="cuda")
a1 = 1.6 * torch.rand(B * C, K, device="cuda") - 0.8
a = torch.stack([torch.ones_like(a1), a1, 0.3 + 0.3 * torch.rand_like(a1)], -1)
b = torch.randn(B * C, K, 3, device="cuda")


def cascade():
    y = x.view(1, B * C, L)
    for k in range(K):
        y = AF.lfilter(y, a[:, k], b[:, k], clamp=False, batching=True)
    return y


ms = timed(cascade)
print(f"cfg2 stock GPU path (5 x torchaudio lfilter, CUDA): {ms:.2f} ms per step -> {B * C * L / ms / 1e6:.3f} Gsamples/s")

# config 3b: 512 x 2 x 131072 with 1023 taps, FFT convolution as the reference writes it
B, C, L, N = 512, 2, 131072, 1023
x = torch.randn(B, C, L, device="cuda")
h = torch.randn(B, C, N, device="cuda")


def fftconv():
    n = L + N - 1
    n += n % 2
    return torch.fft.irfft(torch.fft.rfft(x, n) * torch.fft.rfft(h, n), n)[..., :L]


ms = timed(fftconv)
print(f"cfg3b stock GPU path (torch.fft rfft/irfft, cuFFT): {ms:.2f} ms per step -> {B * C * L / ms / 1e6:.3f} Gsamples/s")
