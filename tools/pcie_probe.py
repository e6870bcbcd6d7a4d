"""Host<->device copy ceiling of the box: pinned 256 MiB H2D alone, D2H alone, both at once (two streams)."""
import time, torch
n = 64 * 1024 * 1024
h_in, h_out = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
d_in, d_out = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
    run(a, b, 2)
    t = run(a, b)
    print(f"{name}: {t*1e3:.2f} ms per 256 MiB each -> {n*4/t/1e9:.1f} GB/s per direction")
