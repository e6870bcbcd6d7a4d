#!/usr/bin/env python
"""bench.py -- throughput of the grafx hot path on B200 (contract: DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5|cfg1|cfg2|cfg2lf|cfg3|cfg3b|cfg4|cfg4b]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

Default workload = BASELINE.json configs[4], the configuration the multi-GPU metric is quoted on: the 32-track mixing
graph (in -> ParametricEqualizer -> Compressor -> STFTMaskedNoiseReverb -> out bus), batch 128, 2 ch x 131072 samples.
The batch of renders is SHARDED over the ranks (128 / N renders per GPU, rendered in chunks of (up to) 32 through
CUDA-graph-captured `render_grafx` plans), the mixes are all-gathered over NCCL (asynchronously, overlapping the next
step) -> `scaling: "strong"`.  At N = 1 the same JSON line also carries `per_config`: one sub-record per other
BASELINE config (cfg1 .. cfg4b) with its own roofline and CPU baseline.  A step = one pass over the whole batch.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio samples/sec (batch x chan x len)"


# --------------------------------------------------------------------------- workloads
class Workload:
    """One processor configuration of BASELINE.json: per-GPU shapes, parameter factory, the module under test,
    algorithmic bytes / flops per step and the oracle call of the CPU arm."""

    def __init__(self, name):
        self.name = name
        self.gen = torch.Generator().manual_seed(0)
        self.cls = None
        if name == "cfg1":
            self.B, self.C, self.L = 1, 2, 48000
            self.desc = "BiquadFilter(num_filters=1, backend=lfilter) 1 x 2ch x 48000 (the reference's CPU-runnable case)"
            self.cls, self.kw = "BiquadFilter", dict(num_filters=1, backend="lfilter", flashfftconv=False)
            self.flops_per_sample = 10.0
        elif name in ("cfg2", "cfg2lf"):
            self.B, self.C, self.L = 256, 2, 131072
            lf = " with every band below 2 kHz (double-precision carry path)" if name == "cfg2lf" else ""
            self.desc = f"ParametricEqualizer(num_filters=5, stereo, backend=lfilter) batch=256 x 2ch x 131072{lf}"
            self.cls, self.kw = "ParametricEqualizer", dict(num_filters=5, processor_channel="stereo", backend="lfilter", flashfftconv=False)
            self.flops_per_sample = 10.0 * 5
        elif name == "cfg3":
            self.B, self.C, self.L = 512, 2, 131072
            self.desc = "STFTMaskedNoiseReverb(ir_len=96000) batch=512 x 2ch x 131072"
            self.cls, self.kw = "STFTMaskedNoiseReverb", dict(ir_len=96000, flashfftconv=False)
            self.flops_per_sample = 8 * math.log2(16384) + 8 * 96000 / 8192
        elif name == "cfg3b":
            self.B, self.C, self.L = 512, 2, 131072
            self.desc = "FIRFilter(fir_len=1023, stereo) batch=512 x 2ch x 131072"
            self.cls, self.kw = "FIRFilter", dict(fir_len=1023, processor_channel="stereo")
            self.flops_per_sample = 8 * math.log2(8192) + 8 * 1023 / (8192 - 1024)
        elif name in ("cfg4", "cfg4b"):
            self.B, self.C, self.L = 1024, 1, 65536
            self.sm = "iir" if name == "cfg4" else "ballistics"
            self.desc = f"SerialChain(Compressor({self.sm}) -> NoiseGate({self.sm})) fused, batch=1024 x 1ch x 65536"
            self.cls, self.kw = "chain", dict(energy_smoother=self.sm, flashfftconv=False)
            self.flops_per_sample = 30.0 * 2
        else:
            raise SystemExit(f"unknown workload {name}")

    # ---- product modules (CUDA)
    def build(self, device):
        import grafx_b200.processors as P

        if self.cls == "chain":
            self.mod = P.SerialChain({"comp": P.Compressor(**self.kw), "gate": P.NoiseGate(**self.kw)}).to(device)
        else:
            self.mod = getattr(P, self.cls)(**self.kw).to(device)

    def param_shapes(self):
        if self.cls == "chain":
            n = 1 if self.sm == "iir" else 2
            one = {"log_threshold": (1,), "log_ratio": (1,), "log_knee": (1,), "z_alpha_pre": (n,)}
            return {"comp": one, "gate": one}
        if self.name == "cfg1":
            return {"Bs": (1, 3), "A1_pre": (1,), "A2_pre": (1,)}
        if self.name in ("cfg2", "cfg2lf"):
            return {k: (2, 5) for k in ("w0", "q_inv", "log_gain")}
        if self.name == "cfg3":
            return {"init_log_magnitude": (2, 193), "delta_log_magnitude": (2, 193)}
        return {"fir": (2, 1023)}

    def host_inputs(self, B=None):
        B = self.B if B is None else B
        x = torch.randn(B, self.C, self.L, generator=self.gen)

        def draw(shapes):
            return {k: (draw(s) if isinstance(s, dict) else torch.randn(B, *s, generator=self.gen)) for k, s in shapes.items()}

        prm = draw(self.param_shapes())
        if self.name == "cfg2lf":
            # w0 = pi * sigmoid(p): p <= -2.4 puts every band below 2 kHz at 48 kHz (down to a few tens of Hz)
            prm["w0"] = -2.4 - 1.5 * torch.rand(B, 2, 5, generator=self.gen) ** 2 * 3.0
        return x, prm

    def forward(self, x, prm):
        out = self.mod(x, **prm)
        return out[0] if isinstance(out, tuple) else out

    def dominant_call(self, x, prm):
        """(callable, kernel name) of the O(samples) kernel(s) without the parameter-side statement around them."""
        import grafx_b200.functional as F_
        from grafx_b200.processors import design

        if self.name in ("cfg2", "cfg2lf"):
            Bs, As = design.parametric_eq(prm["w0"], prm["q_inv"], prm["log_gain"])
            return (lambda: F_.biquad_cascade(x, Bs, As)), "biquad_cascade_x2_kernel (+ cascade_x2_tables_kernel)"
        if self.name == "cfg3":
            return (lambda: self.mod(x, **prm)), "reverb pipeline: reverb_ir + filter spectra + partitioned overlap-save (fir.cu)"
        if self.name == "cfg3b":
            h = F_.normalize_impulse(torch.tanh(prm["fir"]))
            return (lambda: F_.fir_conv(x, h)), "fir_ols_kernel<4096> (+ fir_spectrum_kernel)"
        if self.name == "cfg1":
            return (lambda: self.forward(x, prm)), "biquad_cascade_x2_kernel (+ design, tables)"
        return (lambda: self.forward(x, prm)), "dynamics_kernel (Compressor -> NoiseGate fused, one launch)"

    def alg_bytes(self):
        n = self.B * self.C * self.L
        return 8 * n + (4 * self.B * 2 * 1023 if self.name == "cfg3b" else 0)

    def samples(self, B=None):
        return (self.B if B is None else B) * self.C * self.L

    # ---- CPU arm: the oracle port (torch CPU ops + torchaudio lfilter, the library calls the reference makes)
    def cpu_forward(self, x, prm):
        from oracle import grafx_oracle as O

        if self.name == "cfg1":
            return O.biquad_filter(x, **prm, backend="lfilter")
        if self.name in ("cfg2", "cfg2lf"):
            return O.parametric_equalizer(x, **prm, processor_channel="stereo", backend="lfilter")
        if self.name == "cfg3":
            return O.stft_masked_noise_reverb(x, **prm, ir_len=96000)
        if self.name == "cfg3b":
            return O.fir_filter(x, prm["fir"], "stereo")
        return O.noisegate(O.compressor(x, **prm["comp"], energy_smoother=self.sm), **prm["gate"], energy_smoother=self.sm)

    def cpu_sample_batch(self):
        """Batch of the bounded CPU sample (a few hundred ms to a few s of host work per call)."""
        return {"cfg1": 1, "cfg2": 32, "cfg2lf": 32, "cfg3": 8, "cfg3b": 16, "cfg4": 64, "cfg4b": 16}[self.name]


class GraphWorkload:
    """BASELINE config 5: 32 x (in -> eq -> compressor -> reverb) -> out, batch 128 sharded over the ranks."""

    name = "cfg5"
    # CHUNK: renders per captured plan (every tensor of a chunk stays below 2^31 elements: the signal buffer of 32 renders
    # has 1.08e9); measured on one B200: 16 renders per plan 16.47 ms per step, 32 renders 15.96 ms (profiles/r02_chunk_size.txt)
    TOTAL, CHUNK, TRACKS, C, L = 128, 32, 32, 2, 131072

    def __init__(self, world=1, rank=0, chunk=None, chunk_streams=1):
        self.world, self.rank = world, rank
        self.chunk_streams = max(1, int(chunk_streams))
        self.CHUNK = int(chunk) if chunk else self.CHUNK
        assert self.TOTAL % world == 0, "128 renders split evenly over 1/2/4/8 ranks"
        self.B_local = self.TOTAL // world
        self.chunk = min(self.CHUNK, self.B_local)
        assert self.B_local % self.chunk == 0, "the renders of a rank split evenly into chunks"
        self.n_chunks = self.B_local // self.chunk
        self.desc = ("mixing graph 32 x (in -> ParametricEqualizer(5, stereo, lfilter) -> Compressor -> STFTMaskedNoiseReverb(96000)) "
                     "-> out, batch=128 renders x 2ch x 131072")
        # per node-sample: cascade 10 K + dynamics 30 + reverb (FFT overlap-save) ; + 1 add per track at the bus
        self.flops_per_out_sample = self.TRACKS * (50.0 + 30.0 + 8 * math.log2(16384) + 8 * 96000 / 8192 + 1.0)

    def build(self, device):
        import grafx_b200.processors as P
        from grafx_b200.render import mixing_console_plan

        self.device = device
        self.procs = {"eq": P.ParametricEqualizer(num_filters=5, processor_channel="stereo", backend="lfilter").to(device),
                      "compressor": P.Compressor().to(device), "reverb": P.STFTMaskedNoiseReverb(ir_len=96000).to(device)}
        self.plan = mixing_console_plan(self.TRACKS, ["eq", "compressor", "reverb"])

    def host_params(self, procs=None):
        g = torch.Generator().manual_seed(0)
        procs = procs or self.procs
        return {t: {k: 0.3 * torch.randn(self.TRACKS, *((v,) if isinstance(v, int) else v), generator=g)
                    for k, v in p.parameter_size().items()} for t, p in procs.items()}

    def capture(self):
        """One CUDA-graph-captured render plan per chunk of (up to) 32 renders; the chunk's sources are the capture's static
        input tensor (filled on the device: inputs are HBM-resident when the timed region starts)."""
        from grafx_b200.render import CapturedRender

        self.prm = tree_to(self.host_params(), self.device)
        g = torch.Generator(device=self.device).manual_seed(1000 + self.rank)
        self.caps = []
        for _ in range(self.n_chunks):
            x = torch.randn(self.chunk, self.TRACKS, self.C, self.L, device=self.device, generator=g)
            self.caps.append(CapturedRender(self.procs, x, self.prm, self.plan))
            del x
        # mixes of this rank, double-buffered (the gather of step k reads one while step k+1 fills the other)
        self.mix = [torch.empty(self.B_local, 1, self.C, self.L, device=self.device) for _ in range(2)]
        self.gathered = [torch.empty(self.TOTAL, 1, self.C, self.L, device=self.device) for _ in range(2)] if self.world > 1 else None
        self.pending = None
        self.k = 0

    def step(self):
        """Renders this rank's 128 / N items and starts the all-gather of the mixes (N > 1)."""
        import torch.distributed as dist

        mix = self.mix[self.k & 1]
        if self.chunk_streams > 1 and len(self.caps) > 1:
            # independent chunks on a few streams (experiment: --chunk-streams): fork from / join into the current stream
            main = torch.cuda.current_stream()
            if not hasattr(self, "_side"):
                self._side = [torch.cuda.Stream(self.device) for _ in range(min(self.chunk_streams, len(self.caps)))]
            for st in self._side:
                st.wait_stream(main)
            for i, cap in enumerate(self.caps):
                with torch.cuda.stream(self._side[i % len(self._side)]):
                    out, _, _ = cap()
                    mix[i * self.chunk:(i + 1) * self.chunk].copy_(out)
            for st in self._side:
                main.wait_stream(st)
        else:
            for i, cap in enumerate(self.caps):
                out, _, _ = cap()
                mix[i * self.chunk:(i + 1) * self.chunk].copy_(out)
        if self.world > 1:
            if self.pending is not None:
                self.pending.wait()
            self.pending = dist.all_gather_into_tensor(self.gathered[self.k & 1], mix, async_op=True)
        self.k += 1
        return mix

    def finish(self):
        if self.pending is not None:
            self.pending.wait()
            self.pending = None

    def samples(self):
        return self.TOTAL * self.C * self.L      # rendered output samples per step (whole job)

    def alg_bytes_local(self):
        # contract-preserving render (signal buffer returned): 129 node signals written, 160 read
        return 289 * self.B_local * self.C * self.L * 4

    def cpu_forward(self, x, prm):
        from oracle import grafx_oracle as O

        procs = {"eq": lambda s, **p: O.parametric_equalizer(s, **p, processor_channel="stereo", backend="lfilter"),
                 "compressor": lambda s, **p: O.compressor(s, **p),
                 "reverb": lambda s, **p: O.stft_masked_noise_reverb(s, **p, ir_len=96000)}
        T = self.TRACKS
        plan = {"num_nodes": 4 * T + 1, "iters": [None] + [
            {"type": t, "reads": [("slice", (i * T, (i + 1) * T))], "aggs": [("none", None)], "param": ("slice", (0, T)),
             "write": ("slice", ((i + 1) * T, (i + 2) * T))} for i, t in enumerate(["eq", "compressor", "reverb"])] + [
            {"type": "out", "reads": [("slice", (3 * T, 4 * T))], "aggs": [("sum", None)], "param": ("slice", (0, 1)),
             "write": ("slice", (4 * T, 4 * T + 1))}]}
        return O.render_plan(procs, x, prm, plan)[0]


def config_of(name, world=1, chunk=None):
    """The `config` object of the JSON line -- the SAME dictionary in both arms (the reference arm runs a bounded
    sample of exactly this configuration; what the sample was is said in its `cpu_baseline.sample`)."""
    if name == "cfg5":
        g = GraphWorkload(world, chunk=chunk)
        return {"workload": g.desc, "batch": 128, "tracks": 32, "channels": 2, "length": 131072,
                "renders_per_gpu": g.B_local, "chunks_per_gpu": g.n_chunks,
                "l2": (f"inputs larger than L2: each chunk of {g.chunk} renders reads {g.chunk * 33.554432:.0f} MB of sources and fills a "
                       f"{g.chunk * 0.135266304:.2f} GB signal buffer (126 MiB L2)"),
                "parallelism": (f"batch-of-renders shard x{world}; async NCCL all_gather of the mixes overlapping the next step"
                                if world > 1 else "single GPU: all 128 renders, no collective")}
    wl = Workload(name)
    big = wl.samples() * 4 > 126 * 2**20
    return {"workload": wl.desc, "batch": wl.B, "channels": wl.C, "length": wl.L,
            "l2": (f"inputs larger than L2 ({wl.samples() * 4 / 2**20:.0f} MiB read + as much written per step vs 126 MiB L2)" if big
                   else "working set fits L2 (single latency-bound item, the reference's CPU-runnable case)"),
            "parallelism": f"replicas x{world}: every rank processes its own batch of this size, no collective on the data path"}


def tree_to(obj, device, non_blocking=False):
    if isinstance(obj, torch.Tensor):
        return obj.to(device, non_blocking=non_blocking)
    return {k: tree_to(v, device, non_blocking) for k, v in obj.items()}


def tree_pin(obj):
    if isinstance(obj, torch.Tensor):
        return obj.pin_memory()
    return {k: tree_pin(v) for k, v in obj.items()}


def tree_bytes(obj):
    if isinstance(obj, torch.Tensor):
        return obj.numel() * obj.element_size()
    return sum(tree_bytes(v) for v in obj.values())


def tree_slice(obj, lo, hi):
    if isinstance(obj, torch.Tensor):
        return obj[lo:hi]
    return {k: tree_slice(v, lo, hi) for k, v in obj.items()}


def gpu_local_cpus(gpu_index):
    """CPUs on the NUMA node of the GPU (NVML), or None: pinned staging buffers are allocated from a thread bound to
    them, so that the copy engines do not cross the socket interconnect."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


class pinned_near:
    """Context: allocate pinned host memory from CPUs local to the GPU."""

    def __init__(self, gpu_index):
        self.cpus = gpu_local_cpus(gpu_index)

    def __enter__(self):
        self.saved = os.sched_getaffinity(0)
        if self.cpus:
            os.sched_setaffinity(0, self.cpus)

    def __exit__(self, *exc):
        os.sched_setaffinity(0, self.saved)
        return False


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, busy, mx, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
                if float(f[4]) >= 10.0:
                    busy.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        use = sorted(busy if busy else sm)
        return {"sm_mhz": use[len(use) // 2] if use else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(busy),
                "window": "every timed region of this run: main workload, kernel-only loops, end-to-end loop, per-config records (50 ms period)"}


def pick_cpu_threads(fn):
    """The reference's torch/torchaudio CPU ops do not scale to every core of a many-core host
    (over-subscription); give the CPU arm its best thread count among a few candidates."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    best_t, best_n = 1e30, cands[0]
    for n in cands:
        torch.set_num_threads(n)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best_t, best_n = dt, n
    torch.set_num_threads(best_n)
    return best_n, cores


def cpu_arm(name, steps=None, warmup=1, budget_s=10.0):
    """Times the oracle port on a bounded sample of workload `name` on the host cores.  steps=None: best of as many
    repetitions as fit in `budget_s` (>= 2); otherwise exactly `steps` timed steps after `warmup`."""
    if name == "cfg5":
        wl = GraphWorkload()
        import grafx_b200.processors as P  # parameter_size() only (no CUDA)

        procs = {"eq": P.ParametricEqualizer(num_filters=5, processor_channel="stereo", backend="lfilter"),
                 "compressor": P.Compressor(), "reverb": P.STFTMaskedNoiseReverb(ir_len=96000)}
        prm = wl.host_params(procs)
        Bs = 1
        x = torch.randn(Bs, wl.TRACKS, wl.C, wl.L, generator=torch.Generator().manual_seed(1))
        samples = Bs * wl.C * wl.L
        what = f"cfg5: {Bs} render of the 32-track graph (of 128) x 2ch x 131072"
    else:
        wl = Workload(name)
        Bs = wl.cpu_sample_batch()
        x, prm = wl.host_inputs(Bs)
        samples = wl.samples(Bs)
        what = f"{name}: batch {Bs} (of {wl.B}) x {wl.C}ch x {wl.L}"
    with torch.no_grad():
        threads, cores = pick_cpu_threads(lambda: wl.cpu_forward(x, prm))
        if steps is None:
            best, reps, t_start = 1e30, 0, time.perf_counter()
            while reps < 2 or (time.perf_counter() - t_start < budget_s and reps < 10):
                t0 = time.perf_counter()
                wl.cpu_forward(x, prm)
                best = min(best, time.perf_counter() - t0)
                reps += 1
            dt, how = best, f"best of {reps}"
        else:
            for _ in range(warmup):
                wl.cpu_forward(x, prm)
            t0 = time.perf_counter()
            for _ in range(steps):
                wl.cpu_forward(x, prm)
            dt, how = (time.perf_counter() - t0) / steps, f"mean of {steps} steps after {warmup} warm-up"
    return {"value": samples / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{what}, {how} at the best thread count ({threads} of {cores} host cores); oracle port = torch CPU ops + "
                      "torchaudio lfilter as the reference calls them (the reference is Python: no oracle/_ref binary)",
            "ms_per_sample_step": dt * 1e3}


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    if rank != 0:
        return
    cb = cpu_arm(args.workload, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_sample_step"], "higher_is_better": True,
            "scaling": "strong" if args.workload == "cfg5" else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(args.workload, args.gpus, args.chunk), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- device-side measurement helpers
def capture_step(fn, device):
    """CUDA-graph capture of `fn` (host launch cost out of the timed region).  Returns (replay, mode, launches of the
    library per call, keepalive)."""
    from grafx_b200 import _cabi

    L_ = _cabi.lib()
    cur = torch.cuda.current_stream(device)
    side = torch.cuda.Stream(device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    cur.wait_stream(side)
    torch.cuda.synchronize(device)
    n0 = L_.gfx_kernel_launch_count()
    fn()
    torch.cuda.synchronize(device)
    launches = int(L_.gfx_kernel_launch_count() - n0)
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        return g.replay, "cuda_graph", launches, (g, out)
    except Exception as e:  # a forward that cannot be captured is timed eagerly (and says so)
        torch.cuda.synchronize(device)
        return fn, f"eager ({type(e).__name__})", launches, None


def timed_blocks(step, steps, min_seconds, barrier, device, max_over_ranks, finish=None, max_blocks=400):
    """Blocks of exactly `steps` steps, each bracketed by barrier + synchronize and timed with CUDA events on the
    launching stream; repeated until `min_seconds` of device time (first block decides how many).  Returns
    (ms per step = total over all blocks / steps run, max over ranks; list of block times)."""
    def block():
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        if finish is not None:
            finish()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), device)

    first = block()
    n = int(min(max_blocks, max(1, math.ceil(min_seconds * 1e3 / max(first, 1e-3)))))
    times = [block() for _ in range(n)]
    return sum(times) / (len(times) * steps), times


def fma_peak(device):
    """Measured fp32 FMA rate of this GPU (FMA/s), gfx_fma_probe_f32 timed with events; best of 3."""
    from grafx_b200 import _cabi

    L_ = _cabi.lib()
    out = torch.zeros(16, device=device)
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = L_.gfx_fma_probe_f32(out.data_ptr(), 4096, _cabi.stream_ptr())
        e1.record()
        torch.cuda.synchronize(device)
        if n > 0:
            best = max(best, n / (e0.elapsed_time(e1) * 1e-3))
    return best


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "of measured"
    except Exception:
        return 6650.0, "of fallback"


def traffic_of(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
    except Exception:
        return None


def roofline_of(name, alg_bytes, flops, k_ms, kernel, launches, fma_rate):
    peak, how = hbm_peak()
    ach = alg_bytes / (k_ms * 1e-3) / 1e9
    r = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic_of(name),
         "kernel": kernel, "kernel_ms": k_ms, "launches_per_step": launches, "algorithmic_bytes": alg_bytes, "peak_source": how,
         "algorithmic_flops": flops}
    if fma_rate:
        r["fp32_fma_frac"] = flops / (k_ms * 1e-3) / (2.0 * fma_rate)
        r["fp32_fma_peak_tflops"] = 2.0 * fma_rate / 1e12
    return r


def measure_processor(name, device, steps, warmup, min_seconds, world, barrier, max_over_ranks, fma_rate, local_rank,
                      want_e2e=True, e2e_chunks=4, e2e_streams=3):
    """Device-resident throughput, kernel-only roofline and (optionally) the end-to-end figure of one processor
    workload.  Returns a record with the bench-line fields."""
    wl = Workload(name)
    wl.build(device)
    x_h, prm_h = wl.host_inputs()
    x, prm = x_h.to(device), tree_to(prm_h, device)
    with torch.no_grad():
        step, mode, launches, keep = capture_step(lambda: wl.forward(x, prm), device)
        for _ in range(max(warmup, 3)):
            step()
        ms_step, blocks = timed_blocks(step, steps, min_seconds, barrier, device, max_over_ranks)
        dom_fn, dom_name = wl.dominant_call(x, prm)
        kstep, kmode, klaunches, kkeep = capture_step(dom_fn, device)
        for _ in range(3):
            kstep()
        k_ms, _ = timed_blocks(kstep, steps, min(min_seconds, 0.3), barrier, device, lambda v, d: v)
        rec = {"workload": wl.desc, "value": wl.samples() * world / (ms_step * 1e-3), "unit": "samples/s", "ms_per_step": ms_step,
               "timed_blocks": len(blocks), "steps_per_block": steps, "launch": mode, "launches_per_step": launches,
               "roofline": roofline_of(name, wl.alg_bytes(), wl.flops_per_sample * wl.samples(), k_ms, dom_name, klaunches, fma_rate)}
        del keep, kkeep
        if want_e2e:
            rec["e2e"] = e2e_processor(wl, x_h, prm_h, device, steps, world, barrier, max_over_ranks, local_rank, e2e_chunks, e2e_streams)
    del x, prm
    torch.cuda.empty_cache()
    return rec, wl


def e2e_processor(wl, x_h, prm_h, device, steps, world, barrier, max_over_ranks, local_rank, n_chunks, n_streams):
    """The same metric through the public nn.Module API from pinned HOST tensors: H2D of the inputs, forward, D2H of
    the output audio, chunk-pipelined over streams and software-pipelined over steps."""
    with pinned_near(local_rank):
        xp, prm_p = x_h.pin_memory(), tree_pin(prm_h)
        out_bufs = [torch.empty(wl.B, wl.C, wl.L, dtype=torch.float32).pin_memory() for _ in range(2)]
    nchunk = min(n_chunks, wl.B) if wl.B >= 8 else 1
    streams = [torch.cuda.Stream(device) for _ in range(min(n_streams, nchunk))]
    bounds = [(wl.B * i // nchunk, wl.B * (i + 1) // nchunk) for i in range(nchunk)]

    def enqueue(k):
        out_k = out_bufs[k % 2]
        for i, (lo, hi) in enumerate(bounds):
            s = streams[(k * nchunk + i) % len(streams)]
            with torch.cuda.stream(s):
                xd = xp[lo:hi].to(device, non_blocking=True)
                pd = tree_to(tree_slice(prm_p, lo, hi), device, True)
                out_k[lo:hi].copy_(wl.forward(xd, pd), non_blocking=True)
        evs = []
        for s in streams:
            e = torch.cuda.Event()
            e.record(s)
            evs.append(e)
        return evs

    def run(n):
        pending = None
        for k in range(n):
            evs = enqueue(k)
            if pending is not None:
                for e in pending:
                    e.synchronize()
            pending = evs
        for e in pending:
            e.synchronize()

    run(2)
    barrier()
    n_e2e = max(3, min(steps, 10))
    t0 = time.perf_counter()
    run(n_e2e)
    barrier()
    e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / n_e2e, device)
    return {"value": wl.samples() * world / (e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": tree_bytes(xp) + tree_bytes(prm_p),
            "d2h_bytes_per_step": tree_bytes(out_bufs[0]), "ms_per_step": e_ms,
            "how": f"pinned host tensors -> {nchunk} chunks over {len(streams)} streams: H2D, nn.Module forward, D2H of the output audio; "
                   "steps pipelined (host waits for step k-1 after enqueuing step k, two result buffers)"}


def e2e_graph(wl, device, steps, barrier, max_over_ranks, local_rank):
    """Config 5 end to end: every chunk's sources come from pinned host memory straight into the captured plan's
    static input tensor (H2D stream), the plan is replayed (compute stream), the mixes go back to pinned host memory
    (D2H stream); chunks and steps overlap through events.  Also measures the bare H2D ceiling of the same buffers."""
    with pinned_near(local_rank):
        x_host = [torch.empty(wl.chunk, wl.TRACKS, wl.C, wl.L, dtype=torch.float32).pin_memory() for _ in wl.caps]
        prm_host = tree_pin(wl.host_params())
        out_host = [torch.empty(wl.B_local, 1, wl.C, wl.L, dtype=torch.float32).pin_memory() for _ in range(2)]
    for h, cap in zip(x_host, wl.caps):
        h.copy_(cap.input_signals)  # (same synthetic audio as the device-resident run)
    torch.cuda.synchronize(device)
    s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    cur = torch.cuda.current_stream(device)
    done = [None] * len(wl.caps)      # compute of chunk i (previous step) finished: its input may be overwritten
    drained = [None] * len(wl.caps)   # D2H of chunk i (previous step) finished: its output may be overwritten

    def step(k):
        out_k = out_host[k & 1]
        last = None
        for i, cap in enumerate(wl.caps):
            with torch.cuda.stream(s_in):
                if done[i] is not None:
                    s_in.wait_event(done[i])
                cap.input_signals.copy_(x_host[i], non_blocking=True)
                if i == 0:
                    for t, d in prm_host.items():
                        for n, v in d.items():
                            wl.prm[t][n].copy_(v, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            cur.wait_event(ev_in)
            if drained[i] is not None:
                cur.wait_event(drained[i])
            out, _, _ = cap(None, wl.prm if i == 0 else None)
            done[i] = torch.cuda.Event()
            done[i].record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done[i])
                out_k[i * wl.chunk:(i + 1) * wl.chunk].copy_(out, non_blocking=True)
                drained[i] = torch.cuda.Event()
                drained[i].record(s_out)
                last = drained[i]
        return last

    def run(n):
        pending = None
        for k in range(n):
            ev = step(k)
            if pending is not None:
                pending.synchronize()
            pending = ev
        pending.synchronize()

    run(1)
    barrier()
    n_e2e = max(2, min(steps, 5))
    t0 = time.perf_counter()
    run(n_e2e)
    barrier()
    e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / n_e2e, device)
    # ceiling: the H2D copies alone, all ranks at once (what the host's PCIe / memory path sustains)
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        for h, cap in zip(x_host, wl.caps):
            cap.input_signals.copy_(h, non_blocking=True)
    torch.cuda.synchronize(device)
    barrier()
    c_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / 2, device)
    h2d = sum(tree_bytes(h) for h in x_host) + tree_bytes(prm_host)
    return {"value": wl.samples() / (e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": tree_bytes(out_host[0]), "ms_per_step": e_ms,
            "ceiling_ms_per_step": c_ms, "ceiling_gbs_per_rank": h2d / (c_ms * 1e-3) / 1e9,
            "how": "per rank: pinned host sources -> H2D stream straight into each captured plan's static input, CapturedRender replay, "
                   "mixes D2H on a third stream; chunks and steps overlap through events; `ceiling_*` = the same H2D copies "
                   "alone on all ranks at once (the step is bound by the host-to-device path: 32 input tracks per output mix)"}


# --------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--min-seconds", type=float, default=0.5, help="device time per workload (blocks of --steps steps are repeated)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-per-config", action="store_true")
    ap.add_argument("--per-config", default="cfg1,cfg2,cfg2lf,cfg3,cfg3b,cfg4,cfg4b")
    ap.add_argument("--chunk", type=int, default=None, help="config 5: renders per captured plan (default: GraphWorkload.CHUNK)")
    ap.add_argument("--chunk-streams", type=int, default=1, help="config 5: independent chunks replayed on this many streams")
    ap.add_argument("--e2e-chunks", type=int, default=4)
    ap.add_argument("--e2e-streams", type=int, default=3)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the host arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from grafx_b200.render.parallel import max_over_ranks

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    fma_rate = fma_peak(device)
    steps, warmup = args.steps, max(args.warmup, 3)

    if args.workload == "cfg5":
        wl = GraphWorkload(world, rank, chunk=args.chunk, chunk_streams=args.chunk_streams)
        wl.build(device)
        with torch.no_grad():
            wl.capture()
            for _ in range(warmup):
                wl.step()
            wl.finish()
            ms_step, blocks = timed_blocks(wl.step, steps, args.min_seconds, barrier, device, max_over_ranks, finish=wl.finish)
            value = wl.samples() / (ms_step * 1e-3)
            # the renders alone (no gather): roofline of the graph path on this rank
            def renders_only():
                for cap in wl.caps:
                    cap()
            for _ in range(2):
                renders_only()
            k_ms, _ = timed_blocks(renders_only, max(2, steps // 4), min(args.min_seconds, 0.3), barrier, device, max_over_ranks)
            launches = wl.caps[0].launches_per_replay
            roofline = roofline_of("cfg5", wl.alg_bytes_local(), wl.flops_per_out_sample * wl.B_local * wl.C * wl.L, k_ms,
                                   f"render_grafx, 5 render orders per chunk of {wl.chunk} renders: biquad_cascade_x2 (reads the sources, fills the buffer's source slice), dynamics, "
                                   "reverb pipeline (reverb_ir + filter spectra + partitioned overlap-save), node_sum; CUDA-graph replay",
                                   launches, fma_rate)
            roofline["bytes_convention"] = "contract-preserving render: 289 node-signal passes (129 written, 160 read) x 1 MiB per render"
            roofline["gather_ms_per_step"] = max(ms_step - k_ms, 0.0)
            e2e = None if args.no_e2e else e2e_graph(wl, device, steps, barrier, max_over_ranks, local_rank)
        gpu_launches = None if launches is None else launches * wl.n_chunks * steps * len(blocks) * world
        config = config_of("cfg5", world, args.chunk)
        scaling = "strong"
        main_wl_for_cpu = "cfg5"
        extra = {"timed_blocks": len(blocks), "steps_per_block": steps, "launch": f"cuda_graph (CapturedRender, one plan per chunk of {wl.chunk} renders)"}
    else:
        with torch.no_grad():
            rec, pw = measure_processor(args.workload, device, steps, warmup, args.min_seconds, world, barrier, max_over_ranks,
                                        fma_rate, local_rank, want_e2e=not args.no_e2e, e2e_chunks=args.e2e_chunks,
                                        e2e_streams=args.e2e_streams)
        value, ms_step, roofline, e2e = rec["value"], rec["ms_per_step"], rec["roofline"], rec.get("e2e")
        gpu_launches = rec["launches_per_step"] * steps * rec["timed_blocks"] * world
        config = config_of(args.workload, world)
        scaling = "weak"
        main_wl_for_cpu = args.workload
        extra = {"timed_blocks": rec["timed_blocks"], "steps_per_block": steps, "launch": rec["launch"]}

    # ---- the other BASELINE configs (N = 1 only): one sub-record each
    per_config = None
    if world == 1 and not args.no_per_config and args.workload == "cfg5":
        if args.workload == "cfg5":
            del wl.caps, wl.mix
            torch.cuda.empty_cache()
        per_config = {}
        for name in [n for n in args.per_config.split(",") if n]:
            with torch.no_grad():
                rec, _ = measure_processor(name, device, steps, 3, min(args.min_seconds, 0.5), 1, barrier, max_over_ranks, fma_rate,
                                           local_rank, want_e2e=False)
            per_config[name] = rec
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_arm(main_wl_for_cpu, budget_s=12.0)
        if per_config:
            for name, rec in per_config.items():
                rec["cpu_baseline"] = cpu_arm(name, budget_s=3.0)
                rec["gpu_over_cpu"] = rec["value"] / rec["cpu_baseline"]["value"]

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches,
                "roofline": roofline, "cpu_baseline": cpu_baseline, **extra}
        if per_config is not None:
            line["per_config"] = per_config
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
