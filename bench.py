#!/usr/bin/env python
"""bench.py -- throughput of the grafx hot path on B200 (contract: see README / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg3b|cfg4|cfg4b|cfg5]
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

A step = one pass of the hot path over one batch of synthetic 48 kHz audio.  Default workload =
BASELINE.json configs[1]: ParametricEqualizer, 5 biquads, stereo, batch 256, 131072 samples
(backend "lfilter": the exact cascade).  Under torchrun (N > 1) every rank processes its own
batch of the same size (weak scaling; the path shards over the batch axis with no collective).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_BYTES = 126 * 2**20


# --------------------------------------------------------------------------- workloads
class Workload:
    """name, per-GPU shapes, parameter factory and the module(s) under test."""

    def __init__(self, name, batch_div=1):
        self.name = name
        g = torch.Generator().manual_seed(0)
        self.gen = g
        if name == "cfg2":
            self.B, self.C, self.L = 256 // batch_div, 2, 131072
            self.desc = "ParametricEqualizer(num_filters=5, stereo, backend=lfilter) batch=256 x 2ch x 131072"
            self.kw = dict(num_filters=5, processor_channel="stereo", backend="lfilter", flashfftconv=False)
            self.cls = "ParametricEqualizer"
            self.param_shapes = {k: (2, 5) for k in ("w0", "q_inv", "log_gain")}
        elif name == "cfg3":
            self.B, self.C, self.L = 512 // batch_div, 2, 131072
            self.desc = "STFTMaskedNoiseReverb(ir_len=96000) batch=512 x 2ch x 131072"
            self.kw = dict(ir_len=96000, flashfftconv=False)
            self.cls = "STFTMaskedNoiseReverb"
            self.param_shapes = {"init_log_magnitude": (2, 193), "delta_log_magnitude": (2, 193)}
        elif name == "cfg3b":
            self.B, self.C, self.L = 512 // batch_div, 2, 131072
            self.desc = "FIRFilter(fir_len=1023, stereo) batch=512 x 2ch x 131072"
            self.kw = dict(fir_len=1023, processor_channel="stereo")
            self.cls = "FIRFilter"
            self.param_shapes = {"fir": (2, 1023)}
        elif name in ("cfg4", "cfg4b"):
            self.B, self.C, self.L = 1024 // batch_div, 1, 65536
            sm = "iir" if name == "cfg4" else "ballistics"
            self.desc = f"SerialChain(Compressor({sm}) -> NoiseGate({sm})) fused, batch=1024 x 1ch x 65536"
            self.kw = dict(energy_smoother=sm, flashfftconv=False)
            self.cls = "chain"
            n = 1 if sm == "iir" else 2
            self.param_shapes = {"log_threshold": (1,), "log_ratio": (1,), "log_knee": (1,), "z_alpha_pre": (n,)}
        elif name == "cfg5":
            self.B, self.C, self.L = 16 // batch_div if batch_div <= 16 else 1, 2, 131072
            self.tracks = 32
            self.desc = "mixing graph 32 x (in->eq->compressor->reverb) -> out, batch=16 per GPU (128 over 8), 2ch x 131072"
            self.cls = "graph"
        else:
            raise SystemExit(f"unknown workload {name}")

    # ---- product modules (CUDA)
    def build(self, device):
        import grafx_b200.processors as P

        if self.cls == "chain":
            self.mod = P.SerialChain({"comp": P.Compressor(**self.kw), "gate": P.NoiseGate(**self.kw)}).to(device)
        elif self.cls == "graph":
            from grafx_b200.render import mixing_console_plan

            self.procs = {"eq": P.ParametricEqualizer(num_filters=5, processor_channel="stereo", backend="lfilter").to(device),
                          "compressor": P.Compressor().to(device),
                          "reverb": P.STFTMaskedNoiseReverb(ir_len=96000).to(device)}
            self.plan = mixing_console_plan(self.tracks, ["eq", "compressor", "reverb"])
        else:
            self.mod = getattr(P, self.cls)(**self.kw).to(device)

    def host_inputs(self, pin=False):
        if self.cls == "graph":
            x = torch.randn(self.B, self.tracks, self.C, self.L, generator=self.gen)
            prm = {t: {k: 0.3 * torch.randn(self.tracks, *((v,) if isinstance(v, int) else v), generator=self.gen)
                       for k, v in p.parameter_size().items()} for t, p in self.procs.items()}
        else:
            x = torch.randn(self.B, self.C, self.L, generator=self.gen)
            if self.cls == "chain":
                prm = {n: {k: torch.randn(self.B, *s, generator=self.gen) for k, s in self.param_shapes.items()} for n in ("comp", "gate")}
            else:
                prm = {k: torch.randn(self.B, *s, generator=self.gen) for k, s in self.param_shapes.items()}
        if pin:
            x = x.pin_memory()
        return x, prm

    def forward(self, x, prm):
        if self.cls == "graph":
            from grafx_b200.render import render_grafx

            return render_grafx(self.procs, x, prm, self.plan, parameters_grad=False)[0]
        if self.cls == "chain":
            return self.mod(x, **prm)[0]
        return self.mod(x, **prm)

    def dominant_call(self, x, prm):
        """(callable, kernel name(s), algorithmic bytes per call) of the O(samples) kernel(s) of the workload,
        without the parameter-side statement around them.  Algorithmic bytes: SURVEY.md section 8(d)."""
        import grafx_b200.functional as F_
        from grafx_b200.processors import design

        n = self.B * self.C * self.L
        if self.name == "cfg2":
            Bs, As = design.parametric_eq(prm["w0"], prm["q_inv"], prm["log_gain"])
            return (lambda: F_.biquad_cascade(x, Bs, As)), "biquad_cascade_x2_kernel<4,0> (+ cascade_tables_kernel)", 8 * n
        if self.name == "cfg3":
            return (lambda: self.mod(x, **prm)), \
                "reverb pipeline: reverb_ir + fir_spectrum<8192> + fir_xspec<8192> + fir_mac<12> + fir_inv<8192>", 8 * n
        if self.name == "cfg3b":
            h = F_.normalize_impulse(torch.tanh(prm["fir"]))
            return (lambda: F_.fir_conv(x, h)), "fir_ols_kernel<4096,256> (+ fir_spectrum_kernel)", 8 * n + 4 * h.numel()
        if self.name in ("cfg4", "cfg4b"):
            return (lambda: self.mod(x, **prm)), "dynamics_kernel (Compressor -> NoiseGate fused, one launch)", 8 * n
        if self.name == "cfg5":
            # contract-preserving render (signal buffer returned): 129 node signals written, 160 read
            return (lambda: self.forward(x, prm)), "render_grafx: 5 orders (copy, cascade, dynamics, reverb pipeline, node_sum)", \
                289 * self.B * self.C * self.L * 4
        raise SystemExit(self.name)

    def samples(self):
        if self.cls == "graph":
            return self.B * self.C * self.L       # rendered output samples per graph render
        return self.B * self.C * self.L

    # ---- reference arm: the oracle port on the host CPU (torch CPU ops + torchaudio lfilter,
    # the same library calls the reference makes)
    def cpu_forward(self, x, prm):
        from oracle import grafx_oracle as O

        if self.name == "cfg2":
            return O.parametric_equalizer(x, **prm, processor_channel="stereo", backend="lfilter")
        if self.name == "cfg3":
            return O.stft_masked_noise_reverb(x, **prm, ir_len=96000)
        if self.name == "cfg3b":
            return O.fir_filter(x, prm["fir"], "stereo")
        if self.name in ("cfg4", "cfg4b"):
            sm = "iir" if self.name == "cfg4" else "ballistics"
            return O.noisegate(O.compressor(x, **prm["comp"], energy_smoother=sm), **prm["gate"], energy_smoother=sm)
        if self.name == "cfg5":
            procs = {"eq": lambda s, **p: O.parametric_equalizer(s, **p, processor_channel="stereo", backend="lfilter"),
                     "compressor": lambda s, **p: O.compressor(s, **p),
                     "reverb": lambda s, **p: O.stft_masked_noise_reverb(s, **p, ir_len=96000)}
            T = self.tracks
            plan = {"num_nodes": 4 * T + 1, "iters": [None] + [
                {"type": t, "reads": [("slice", (i * T, (i + 1) * T))], "aggs": [("none", None)], "param": ("slice", (0, T)),
                 "write": ("slice", ((i + 1) * T, (i + 2) * T))} for i, t in enumerate(["eq", "compressor", "reverb"])] + [
                {"type": "out", "reads": [("slice", (3 * T, 4 * T))], "aggs": [("sum", None)], "param": ("slice", (0, 1)),
                 "write": ("slice", (4 * T, 4 * T + 1))}]}
            return O.render_plan(procs, x, prm, plan)[0]
        raise SystemExit(self.name)


def tree_to(obj, device, non_blocking=False):
    if isinstance(obj, torch.Tensor):
        return obj.to(device, non_blocking=non_blocking)
    return {k: tree_to(v, device, non_blocking) for k, v in obj.items()}


def tree_pin(obj):
    if isinstance(obj, torch.Tensor):
        return obj.pin_memory()
    return {k: tree_pin(v) for k, v in obj.items()}


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


def tree_bytes(obj):
    if isinstance(obj, torch.Tensor):
        return obj.numel() * obj.element_size()
    return sum(tree_bytes(v) for v in obj.values())


def tree_slice(obj, lo, hi):
    if isinstance(obj, torch.Tensor):
        return obj[lo:hi]
    return {k: tree_slice(v, lo, hi) for k, v in obj.items()}


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
                "window": "warm-up + timed steps + kernel-only loop + end-to-end loop + 1 s of back-to-back steps (50 ms period)"}


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


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    if rank != 0:
        return
    wl = Workload(args.workload)
    if wl.cls == "graph":
        wl.B = 1
        import grafx_b200.processors as P  # parameter_size() only (no CUDA)

        wl.procs = {"eq": P.ParametricEqualizer(num_filters=5, processor_channel="stereo", backend="lfilter"),
                    "compressor": P.Compressor(), "reverb": P.STFTMaskedNoiseReverb(ir_len=96000)}
    else:
        # bounded sample of the workload: 1/8 of the batch per step
        wl.B = max(1, wl.B // 8)
    x, prm = wl.host_inputs()
    with torch.no_grad():
        threads, cores = pick_cpu_threads(lambda: wl.cpu_forward(x, prm))
        for _ in range(args.warmup):
            wl.cpu_forward(x, prm)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            wl.cpu_forward(x, prm)
        dt = (time.perf_counter() - t0) / args.steps
    val = wl.samples() / dt
    line = {"impl": "reference", "metric": "audio samples/sec (batch x chan x len)", "value": val, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.desc, "sample": f"batch {wl.B} per step (bounded sample of the workload)"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{wl.name}: batch {wl.B} x {wl.C}ch x {wl.L}, oracle port (torch CPU ops + torchaudio lfilter as the reference calls them), {args.steps} steps, best of thread counts <= {cores} host cores"},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="batch chunks per end-to-end step (copy/compute overlap)")
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

    batch_div = 1
    if args.workload == "cfg5":
        batch_div = 1  # 16 renders per GPU (= 128 over 8 GPUs)
    wl = Workload(args.workload, batch_div)
    wl.build(device)
    x_h, prm_h = wl.host_inputs(pin=False)
    x = x_h.to(device)
    prm = tree_to(prm_h, device)
    samples = wl.samples()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            y = wl.forward(x, prm)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        from grafx_b200 import _cabi
        launches0 = _cabi.lib().gfx_kernel_launch_count()
        ev0.record()
        for _ in range(args.steps):
            y = wl.forward(x, prm)
        ev1.record()
        barrier()
        launches_timed = int(_cabi.lib().gfx_kernel_launch_count() - launches0)
        ms_total = max_over_ranks(ev0.elapsed_time(ev1), device)
        ms_step = ms_total / args.steps
        value = samples * world / (ms_step * 1e-3)

        # ---- dominant kernel(s) alone (device time, events on the launching stream = torch's current stream)
        dom_fn, dom_name, alg = wl.dominant_call(x, prm)
        L_ = _cabi.lib()
        for _ in range(3):
            dom_fn()
        torch.cuda.synchronize()
        n0 = L_.gfx_kernel_launch_count()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(args.steps):
            dom_fn()
        k1.record()
        torch.cuda.synchronize()
        k_ms = k0.elapsed_time(k1) / args.steps
        dom_launches = (L_.gfx_kernel_launch_count() - n0) / args.steps
        peak, how = 6650.0, "of fallback"
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            how = "of measured"
        except Exception:
            pass
        ach = alg / (k_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl.name)
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "kernel": dom_name, "kernel_ms": k_ms, "launches_per_step": dom_launches,
                    "algorithmic_bytes": alg, "peak_source": how}

        # ---- end to end through the public nn.Module API with HOST buffers (pinned), copies inside
        e2e = None
        if not args.no_e2e:
            saved_affinity = os.sched_getaffinity(0)
            local_cpus = gpu_local_cpus(local_rank)
            if local_cpus:
                os.sched_setaffinity(0, local_cpus)  # (this thread only: the CPU-baseline leg keeps every core)
            try:
                xp, prm_p = x_h.pin_memory(), tree_pin(prm_h)
                out_p = torch.empty((wl.B, 1 if wl.cls == "graph" else wl.C, wl.L) if wl.cls != "graph" else (wl.B, 1, wl.C, wl.L),
                                    dtype=torch.float32).pin_memory()
                out_b = torch.empty_like(out_p).pin_memory()
            finally:
                os.sched_setaffinity(0, saved_affinity)
            nchunk = min(args.e2e_chunks, wl.B) if wl.B >= 8 else 1
            streams = [torch.cuda.Stream(device) for _ in range(min(args.e2e_streams, nchunk))]
            bounds = [(wl.B * i // nchunk, wl.B * (i + 1) // nchunk) for i in range(nchunk)]
            per_item_params = wl.cls != "graph"

            out_bufs = [out_p, out_b]

            def e2e_enqueue(k):
                """Enqueues step k (H2D, forward, D2H of every chunk) and returns the events that mark its results
                as readable on the host.  Results alternate between two pinned buffers, so the host may still be
                reading step k-1 while step k is in flight."""
                out_k = out_bufs[k % 2]
                for i, (lo, hi) in enumerate(bounds):
                    s = streams[(k * nchunk + i) % len(streams)]
                    with torch.cuda.stream(s):
                        xd = xp[lo:hi].to(device, non_blocking=True)
                        pd = tree_to(tree_slice(prm_p, lo, hi) if per_item_params else prm_p, device, True)
                        yd = wl.forward(xd, pd)
                        out_k[lo:hi].copy_(yd, non_blocking=True)
                evs = []
                for s in streams:
                    e = torch.cuda.Event()
                    e.record(s)
                    evs.append(e)
                return evs

            def e2e_run(n):
                # software pipeline over steps: the host waits for step k-1 right after it has enqueued step k, so
                # the copy engines do not idle at step boundaries; every step's result is observed on the host
                pending = None
                for k in range(n):
                    evs = e2e_enqueue(k)
                    if pending is not None:
                        for e in pending:
                            e.synchronize()
                    pending = evs
                for e in pending:
                    e.synchronize()

            e2e_run(2)
            barrier()
            t0 = time.perf_counter()
            n_e2e = max(3, min(args.steps, 10))
            e2e_run(n_e2e)
            barrier()
            e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / n_e2e, device)
            e2e = {"value": samples * world / (e_ms * 1e-3), "unit": "samples/s",
                   "h2d_bytes_per_step": tree_bytes(xp) + (tree_bytes(prm_p)),
                   "d2h_bytes_per_step": tree_bytes(out_p), "ms_per_step": e_ms,
                   "how": f"pinned host tensors -> {nchunk} chunks over {len(streams)} streams: H2D, nn.Module forward, D2H of the output audio; "
                          "steps pipelined (host waits for step k-1 after enqueuing step k, two result buffers)"}

    # the timed region lasts milliseconds -- shorter than one nvidia-smi sample -- so the same step is also run
    # back to back for about a second with the sampler on (not timed): these are the clocks "under load"
    with torch.no_grad():
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            for _ in range(20):
                wl.forward(x, prm)
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline beside it (rank 0, N = 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cwl = Workload(args.workload)
        if cwl.cls == "graph":
            cwl.B = 1
            cwl.procs = wl.procs
        else:
            cwl.B = max(1, cwl.B // 4)
        cx, cprm = cwl.host_inputs()
        with torch.no_grad():
            threads, cores = pick_cpu_threads(lambda: cwl.cpu_forward(cx, cprm))
            best = 1e30
            t_start = time.perf_counter()
            reps = 0
            while reps < 3 or (time.perf_counter() - t_start < 10 and reps < 10):
                t0 = time.perf_counter()
                cwl.cpu_forward(cx, cprm)
                best = min(best, time.perf_counter() - t0)
                reps += 1
        cpu_baseline = {"value": cwl.samples() / best, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{cwl.name}: batch {cwl.B} x {cwl.C}ch x {cwl.L} (1/4 of the workload), best of {reps} at the best thread count ({threads} of {cores} host cores), oracle port = torch CPU ops + torchaudio lfilter as the reference calls them"}

    if rank == 0:
        line = {"metric": "audio samples/sec (batch x chan x len)", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl.desc, "per_gpu_batch": wl.B, "channels": wl.C, "length": wl.L,
                           "l2": f"inputs larger than L2 ({tree_bytes(x) / 2**20:.0f} MiB read + as much written per step vs 126 MiB L2)",
                           "parallelism": f"batch shard x{world}, no collective on the data path"},
                "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches_timed * world,
                "roofline": roofline, "cpu_baseline": cpu_baseline}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
