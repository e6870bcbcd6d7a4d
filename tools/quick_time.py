"""Ad-hoc device timing of single kernels (CUDA events, inputs >> L2).  Not the bench."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import grafx_b200.functional as F_
from grafx_b200.processors import design


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(it)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    torch.manual_seed(0)
    dev = "cuda"
    out = {}
    for (B, C, L, K) in [(256, 2, 131072, 5), (256, 2, 131072, 1), (1024, 1, 65536, 2), (32, 2, 131072, 5), (4, 2, 131072, 5)]:
        x = torch.randn(B, C, L, device=dev)
        w0, q, g = (torch.randn(B, C, K, device=dev) for _ in range(3))
        Bs, As = design.parametric_eq(w0, q, g, use_shelving_filters=K >= 3)
        med, best = timeit(lambda: F_.biquad_cascade(x, Bs, As))
        n = B * C * L
        out[f"cascade_B{B}_C{C}_L{L}_K{K}"] = dict(ms=med, best_ms=best, gsamples_s=n / med / 1e6, gbs=8 * n / med / 1e6)
        print(f"cascade B{B} C{C} L{L} K{K}: {med*1e3:.1f} us (best {best*1e3:.1f})  {n/med/1e6:.1f} Gsamples/s  {8*n/med/1e6:.0f} GB/s", flush=True)
    # the existing sm_100-compiled kernel to beat: torchaudio lfilter on the same GPU
    try:
        from torchaudio.functional import lfilter
        B, C, L, K = 256, 2, 131072, 5
        x = torch.randn(B * C, L, device=dev)
        w0, q, g = (torch.randn(B, C, K, device=dev) for _ in range(3))
        Bs, As = design.parametric_eq(w0, q, g)
        Bf, Af = Bs.reshape(B * C, K, 3), As.reshape(B * C, K, 3)
        def ref():
            y = x
            for i in range(K):
                y = lfilter(y, b_coeffs=Bf[:, i], a_coeffs=Af[:, i], batching=True, clamp=False)
            return y
        med, best = timeit(ref, warm=1, it=3)
        print(f"torchaudio lfilter CUDA (reference GPU path) cfg2: {med:.2f} ms", flush=True)
        out["torchaudio_lfilter_cuda_cfg2_ms"] = med
    except Exception as e:
        print("torchaudio cuda lfilter failed:", e)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/quick_time.json", "w"), indent=1)


if __name__ == "__main__":
    main()
