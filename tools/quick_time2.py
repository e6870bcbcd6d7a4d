"""Ad-hoc device timings of every workload (CUDA events).  Not the bench."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench


def timeit(fn, warm=2, it=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(it)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


out = {}
names = sys.argv[1:] or ["cfg2", "cfg3b", "cfg4", "cfg4b", "cfg3", "cfg5"]
for name in names:
    try:
        wl = bench.Workload(name)
        if name == "cfg5":
            wl.B = 4
        wl.build("cuda")
        x, prm = wl.host_inputs()
        x = x.cuda(); prm = bench.tree_to(prm, "cuda")
        with torch.no_grad():
            ms = timeit(lambda: wl.forward(x, prm))
        n = wl.samples()
        out[name] = dict(ms=ms, gsamples_s=n / ms / 1e6)
        print(f"{name}: {ms:.3f} ms  {n/ms/1e6:.1f} Gsamples/s  (8B/sample -> {8*n/ms/1e6:.0f} GB/s)  [{wl.desc}]", flush=True)
        del x, prm, wl
        torch.cuda.empty_cache()
    except Exception as e:
        import traceback; traceback.print_exc()
        print(name, "FAILED", e, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/quick_time2.json", "w"), indent=1)
