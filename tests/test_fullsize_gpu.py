"""GPU parity at the FULL BASELINE.json sizes ("untested configs"): the product runs the whole
configuration, a sampled set of rows is compared with the oracle on the host; plus the stand-alone envelope modules and
the render_grafx branches that had no test (one-by-one buffer, common_parameters).

Tolerance: rel-L2 <= 1e-4 and max|err| <= 1e-4 * max|ref| per compared block (north_star: within 1e-4 relative)."""
import os

import numpy as np
import pytest
import torch

from _golden import GOLDEN, max_rel, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-4


def assert_close(y, y_ref, name, tol=TOL):
    assert y.shape == y_ref.shape, (name, y.shape, y_ref.shape)
    assert torch.isfinite(y).all(), name
    r, m = rel_l2(y, y_ref), max_rel(y, y_ref)
    assert r <= tol and m <= tol, (name, r, m)


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


# ------------------------------------------------------------------ config 3: reverb, 512 x 2 x 131072, 96000 taps
def test_cfg3_reverb_full_size_rows_vs_oracle():
    from oracle import grafx_oracle as O
    import grafx_b200.processors as P

    gen = torch.Generator().manual_seed(3)
    B, L = 512, 131072
    proc = P.STFTMaskedNoiseReverb(ir_len=96000).cuda()
    x = torch.randn(B, 2, L, generator=gen)
    prm = {k: torch.randn(B, *v, generator=gen) for k, v in proc.parameter_size().items()}
    y = proc(x.cuda(), **_cuda(prm))
    torch.cuda.synchronize()
    rows = [0, 129, 300, 511]
    y_ref = O.stft_masked_noise_reverb(x[rows], **{k: v[rows] for k, v in prm.items()}, ir_len=96000)
    for i, r in enumerate(rows):
        assert_close(y[r].cpu(), y_ref[i], f"cfg3 row {r}")
    # linearity over the whole batch (size-independent property)
    x2 = torch.randn(B, 2, L, generator=gen).cuda()
    lhs = proc(x.cuda() - 0.5 * x2, **_cuda(prm))
    rhs = y - 0.5 * proc(x2, **_cuda(prm))
    assert rel_l2(lhs.cpu(), rhs.cpu()) < 2e-5


# ------------------------------------------------------------------ config 4 / 4b: Compressor -> NoiseGate, 1024 x 1 x 65536
@pytest.mark.parametrize("smoother", ["iir", "ballistics"])
def test_cfg4_chain_full_size_rows_vs_oracle(smoother):
    from oracle import grafx_oracle as O
    import grafx_b200.processors as P

    gen = torch.Generator().manual_seed(44)
    B, L = 1024, 65536
    x = torch.randn(B, 1, L, generator=gen)
    comp, gate = P.Compressor(energy_smoother=smoother), P.NoiseGate(energy_smoother=smoother)

    def draw(proc):
        return {k: torch.randn(B, v, generator=gen) for k, v in proc.parameter_size().items()}

    pc, pg = draw(comp), draw(gate)
    chain = P.SerialChain({"comp": comp, "gate": gate}).cuda()
    y, _ = chain(x.cuda(), comp=_cuda(pc), gate=_cuda(pg))
    torch.cuda.synchronize()
    rows = [0, 1, 257, 600, 1023]
    y1 = O.compressor(x[rows], **{k: v[rows] for k, v in pc.items()}, energy_smoother=smoother)
    y_ref = O.noisegate(y1, **{k: v[rows] for k, v in pg.items()}, energy_smoother=smoother)
    for i, r in enumerate(rows):
        assert_close(y[r].cpu(), y_ref[i], f"cfg4-{smoother} row {r}")


# ------------------------------------------------------------------ config 5: one 32-track render, output AND buffer
def test_cfg5_graph_32_tracks_vs_oracle_render():
    """32 x (in -> eq -> compressor -> reverb) -> out at L = 131072, B = 2 renders: the `filter_repeat` /
    `shared_parameters` path of the product against the oracle's restatement of render_grafx (which has no such path)."""
    from oracle import grafx_oracle as O
    import grafx_b200.processors as P
    from grafx_b200.render import mixing_console_plan, render_grafx

    gen = torch.Generator().manual_seed(5)
    T, B, L = 32, 2, 131072
    procs = {"eq": P.ParametricEqualizer(num_filters=5, processor_channel="stereo", backend="lfilter").cuda(),
             "compressor": P.Compressor().cuda(), "reverb": P.STFTMaskedNoiseReverb(ir_len=96000).cuda()}
    x = torch.randn(B, T, 2, L, generator=gen)
    prm = {t: {k: 0.3 * torch.randn(T, *((v,) if isinstance(v, int) else v), generator=gen)
               for k, v in p.parameter_size().items()} for t, p in procs.items()}
    rd = mixing_console_plan(T, ["eq", "compressor", "reverb"])
    out, inter, buf = render_grafx(procs, x.cuda(), {t: _cuda(p) for t, p in prm.items()}, rd, parameters_grad=False)
    torch.cuda.synchronize()
    oprocs = {"eq": lambda s, **p: O.parametric_equalizer(s, **p, processor_channel="stereo", backend="lfilter"),
              "compressor": lambda s, **p: O.compressor(s, **p),
              "reverb": lambda s, **p: O.stft_masked_noise_reverb(s, **p, ir_len=96000)}
    plan = {"num_nodes": 4 * T + 1, "iters": [None] + [
        {"type": t, "reads": [("slice", (i * T, (i + 1) * T))], "aggs": [("none", None)], "param": ("slice", (0, T)),
         "write": ("slice", ((i + 1) * T, (i + 2) * T))} for i, t in enumerate(["eq", "compressor", "reverb"])] + [
        {"type": "out", "reads": [("slice", (3 * T, 4 * T))], "aggs": [("sum", None)], "param": ("slice", (0, 1)),
         "write": ("slice", (4 * T, 4 * T + 1))}]}
    ref_out, ref_buf = O.render_plan(oprocs, x, prm, plan)
    assert inter == []
    assert_close(out.cpu(), ref_out, "cfg5 output")
    assert tuple(buf.shape) == (B, 4 * T + 1, 2, L)
    # every node signal of the buffer, stage by stage (sources, eq, compressor, reverb, mix)
    for name, lo, hi in (("in", 0, T), ("eq", T, 2 * T), ("compressor", 2 * T, 3 * T), ("reverb", 3 * T, 4 * T), ("out", 4 * T, 4 * T + 1)):
        assert_close(buf[:, lo:hi].cpu(), ref_buf[:, lo:hi], f"cfg5 buffer[{name}]")


# ------------------------------------------------------------------ stand-alone envelope modules
def test_envelope_modules_vs_reference_golden():
    from grafx_b200.processors.core.envelope import Ballistics, TruncatedOnePoleIIRFilter
    from grafx_b200.processors.dynamics import BallisticsEnvelopeFollower, IIREnvelopeFollower

    zf = np.load(os.path.join(GOLDEN, "envelope_modules.npz"))
    u, x = torch.from_numpy(zf["u"]).cuda(), torch.from_numpy(zf["x"]).cuda()
    z1, z2 = torch.from_numpy(zf["z1"]).cuda(), torch.from_numpy(zf["z2"]).cuda()
    n = int(zf["iir_len"])
    assert_close(TruncatedOnePoleIIRFilter(iir_len=n).cuda()(u, z1).cpu(), torch.from_numpy(zf["y_onepole"]), "one-pole")
    assert_close(Ballistics().cuda()(u, z2).cpu(), torch.from_numpy(zf["y_ballistics"]), "ballistics")
    # the followers return log(envelope + 1e-5): where the envelope is ~0 the reference's own FFT-convolution noise
    # (~1e-7 of the envelope's peak) is amplified by the log, so the 1e-4 criterion is applied to envelope + 1e-5
    for det in ("energy", "amplitude"):
        assert_close(IIREnvelopeFollower(detect_with=det, iir_len=n).cuda()(x, z1).cpu().exp(),
                     torch.from_numpy(zf[f"env_iir_{det}"]).exp(), f"iir follower {det}")
        assert_close(BallisticsEnvelopeFollower(detect_with=det).cuda()(x, z2).cpu().exp(),
                     torch.from_numpy(zf[f"env_ballistics_{det}"]).exp(), f"ballistics follower {det}")
    with pytest.raises(ValueError):
        IIREnvelopeFollower(detect_with="rms_channel")


@pytest.mark.parametrize("L", [1, 31, 8193, 70001])
def test_envelope_modules_ragged_vs_oracle(L):
    from oracle import grafx_oracle as O
    from grafx_b200.processors.core.envelope import Ballistics, TruncatedOnePoleIIRFilter
    from grafx_b200.processors.dynamics import IIREnvelopeFollower

    gen = torch.Generator().manual_seed(L)
    B = 5
    u = torch.randn(B, L, generator=gen)  # signed input: the relu of the one-pole matters
    z1 = torch.tensor([[9.0], [2.0], [0.0], [-3.0], [5.0]])
    z2 = torch.randn(B, 2, generator=gen)
    assert_close(TruncatedOnePoleIIRFilter(iir_len=4096).cuda()(u.cuda(), z1.cuda()).cpu(),
                 O.truncated_one_pole_recursive(u, z1, iir_len=4096), f"one-pole L={L}", tol=2e-5)
    assert_close(Ballistics().cuda()(u.abs().cuda(), z2.cuda()).cpu(), O.ballistics(u.abs(), z2), f"ballistics L={L}", tol=2e-5)
    x = torch.randn(B, 3, L, generator=gen)
    got = IIREnvelopeFollower(iir_len=4096).cuda()(x.cuda(), z1.cuda()).cpu()
    ref = torch.log(O.truncated_one_pole_recursive(x.square().mean(-2), z1, iir_len=4096) + 1e-5)
    assert float((got - ref).abs().max()) < 1e-4 * float(ref.abs().max())


def test_ballistics_equal_coefficients_is_one_pole_on_gpu():
    """The reference-side invariant that pins the recursion (tests/test_oracle_golden.py) on the CUDA kernel itself:
    at == rt must reproduce the linear one-pole started from 1, at the full config-4 row length."""
    import scipy.signal
    from grafx_b200.processors.core.envelope import Ballistics

    gen = torch.Generator().manual_seed(9)
    B, L = 6, 65536
    u = torch.rand(B, L, generator=gen) * 2
    z = torch.randn(B, 1, generator=gen).expand(B, 2).contiguous()
    y = Ballistics().cuda()(u.cuda(), z.cuda()).cpu().double().numpy()
    a = torch.sigmoid(z[:, 0].double()).numpy()
    for r in range(B):
        ref, _ = scipy.signal.lfilter([a[r]], [1.0, -(1.0 - a[r])], u[r].double().numpy(), zi=[(1.0 - a[r])])
        assert np.abs(y[r] - ref).max() <= 2e-5 * np.abs(ref).max()


# ------------------------------------------------------------------ render_grafx branches without a test so far
def _three_track_plan(method):
    from grafx_b200.render import mixing_console_plan

    return mixing_console_plan(3, ["eq", "compressor"], method=method)


def test_render_one_by_one_list_buffer_matches_tensor_buffer():
    """`method="one-by-one"`: the signal buffer is a Python list with one entry per node (render/core.py:15-17,81-82);
    one node per render order."""
    import grafx_b200.processors as P
    from grafx_b200.render import render_grafx
    from grafx_b200.render.plan import RenderData, _AggregationData as Agg, _SingleRenderData as It, _TensorAccessData as Acc

    torch.manual_seed(31)
    L = 5000
    x = torch.randn(2, 2, L, device="cuda")
    eq, comp = P.ParametricEqualizer(num_filters=3).cuda(), P.Compressor().cuda()
    prm = {"eq": {k: 0.5 * torch.randn(2, *v, device="cuda") for k, v in eq.parameter_size().items()},
           "compressor": {k: torch.randn(2, v, device="cuda") for k, v in comp.parameter_size().items()}}
    procs = {"eq": eq, "compressor": comp}
    # nodes: 0,1 inputs | 2 = eq(0) | 3 = eq(1) | 4 = compressor(2) | 5 = compressor(3) | 6 = out(4 + 5)
    def one(t, src, par, dst):
        return It(t, [Acc("slice", (src, src + 1))], [Agg("none")], Acc("slice", (par, par + 1)), Acc("slice", (dst, dst + 1)))
    iters = [It("in", [Acc("none", ())], [Agg("none")], Acc("slice", (0, 2)), Acc("slice", (0, 2))),
             one("eq", 0, 0, 2), one("eq", 1, 1, 3), one("compressor", 2, 0, 4), one("compressor", 3, 1, 5)]
    rd_list = RenderData("one-by-one", 6, 4, True, iters)
    out_l, _, buf_l = render_grafx(procs, x, prm, rd_list)
    assert isinstance(buf_l, list) and len(buf_l) == 6
    batched = RenderData("beam", 6, 2, True, [iters[0],
        It("eq", [Acc("slice", (0, 2))], [Agg("none")], Acc("slice", (0, 2)), Acc("slice", (2, 4))),
        It("compressor", [Acc("slice", (2, 4))], [Agg("none")], Acc("slice", (0, 2)), Acc("slice", (4, 6)))])
    out_t, _, buf_t = render_grafx(procs, x, prm, batched)
    for n in range(6):
        assert torch.allclose(buf_l[n][0], buf_t[n], rtol=1e-5, atol=1e-6), n
    assert torch.allclose(out_l[0], buf_t[5], rtol=1e-5, atol=1e-6)


def test_render_common_parameters_reach_the_processor():
    """`common_parameters` (render/graph.py:132-141): per-NODE tensors read with the destination access of every
    render order and passed as extra keyword arguments -- here the DryWet weight of each track."""
    import grafx_b200.processors as P
    from grafx_b200.render import mixing_console_plan, render_grafx

    torch.manual_seed(32)
    T, L = 3, 4000
    x = torch.randn(T, 2, L, device="cuda")
    wet = P.DryWet(P.TanhDistortion()).cuda()
    inner = wet.processor.parameter_size()
    prm = {"fx": {k: 0.5 * torch.randn(T, *((v,) if isinstance(v, int) else v), device="cuda") for k, v in inner.items()}}
    rd = mixing_console_plan(T, ["fx"])
    num_nodes = rd.num_nodes
    weights = torch.rand(num_nodes, 1, device="cuda")
    out, _, buf = render_grafx({"fx": wet}, x, prm, rd, common_parameters={"drywet_weight": weights})
    ref = wet(x, drywet_weight=weights[T:2 * T], **prm["fx"])
    ref = ref[0] if isinstance(ref, tuple) else ref
    assert torch.allclose(buf[T:2 * T], ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(out[0], ref.sum(0), rtol=1e-5, atol=1e-5)
    # a single tensor instead of a dict is passed as `parameter=` upstream; processors without that argument raise
    with pytest.raises(TypeError):
        render_grafx({"fx": wet}, x, prm, rd, common_parameters=weights)


# ------------------------------------------------------------------ ballistics: warm-up chunks vs the row walk
@pytest.mark.parametrize("C,L", [(1, 65536), (2, 40001), (1, 3000), (1, 7), (2, 63), (2, 1)])
def test_ballistics_speculative_chunks_match_the_row_walk(C, L):
    """dynamics_spec_kernel (independent chunks, contraction warm-up) against the sequential row walk of the same
    library (gfx_dynamics_set_ballistics_mode(0)) and the oracle: fast followers, a few slow rows that must fall back
    to the walk inside the same launch, energy AND gain followers, two fused stages."""
    from oracle import grafx_oracle as O
    import grafx_b200.processors as P
    from grafx_b200 import _cabi

    gen = torch.Generator().manual_seed(100 + C)
    B = 24
    x = torch.randn(B, C, L, generator=gen)
    comp = P.Compressor(energy_smoother="ballistics", gain_smoother="ballistics", gain_smooth_in_log=True)
    gate = P.NoiseGate(energy_smoother="ballistics")

    def draw(proc):
        return {k: torch.randn(B, v, generator=gen) for k, v in proc.parameter_size().items()}

    pc, pg = draw(comp), draw(gate)
    pc["z_alpha_pre"][3] = torch.tensor([-7.0, 0.3])    # release-side time constant ~1100 samples: beyond 16 chunks -> walk
    pg["z_alpha_pre"][5] = torch.tensor([1.0, -9.0])
    pc["z_alpha_post"][7] = torch.tensor([-3.0, -3.5])  # slow but inside the warm-up budget
    chain = P.SerialChain({"comp": comp, "gate": gate}).cuda()
    L_ = _cabi.lib()
    try:
        assert L_.gfx_dynamics_set_ballistics_mode(1) == 0
        y_spec, _ = chain(x.cuda(), comp=_cuda(pc), gate=_cuda(pg))
        assert L_.gfx_dynamics_set_ballistics_mode(0) == 0
        y_walk, _ = chain(x.cuda(), comp=_cuda(pc), gate=_cuda(pg))
    finally:
        L_.gfx_dynamics_set_ballistics_mode(1)
    y_spec, y_walk = y_spec.cpu(), y_walk.cpu()
    for r in range(B):
        assert rel_l2(y_spec[r], y_walk[r]) < 2e-6, (r, rel_l2(y_spec[r], y_walk[r]))
    assert torch.equal(y_spec[3], y_walk[3]) and torch.equal(y_spec[5], y_walk[5])   # the rows that fell back
    assert torch.isfinite(y_spec).all()
    rows = [0, 3, 5, 7, 11]
    y1 = O.compressor(x[rows], **{k: v[rows] for k, v in pc.items()}, energy_smoother="ballistics", gain_smoother="ballistics",
                      gain_smooth_in_log=True)
    y_ref = O.noisegate(y1, **{k: v[rows] for k, v in pg.items()}, energy_smoother="ballistics")
    for i, r in enumerate(rows):
        assert_close(y_spec[r], y_ref[i], f"spec row {r}")


def test_channel_broadcast_and_float64_guard():
    """Mono -> stereo broadcasting of StereoGain and the DryWet mix as upstream's tensor expressions do
    (stereo.py:38-41, container.py:62-65); float64 FIR operands fail loudly instead of being computed in float32."""
    import grafx_b200.functional as F_
    import grafx_b200.processors as P

    torch.manual_seed(77)
    x1 = torch.randn(3, 1, 4001, device="cuda")
    lg = torch.randn(3, 2, device="cuda")
    y = P.StereoGain().cuda()(x1, lg)
    assert torch.allclose(y, x1 * torch.exp(lg)[:, :, None], rtol=1e-6, atol=1e-7)
    wet = torch.randn(3, 2, 4001, device="cuda")
    w = torch.rand(3, 1, device="cuda")
    mixed = F_.drywet_mix(x1, wet, w)
    assert torch.allclose(mixed, w.view(-1, 1, 1) * wet + (1 - w.view(-1, 1, 1)) * x1, rtol=1e-6, atol=1e-6)
    with pytest.raises(TypeError):
        F_.fir_conv(wet.double(), torch.randn(3, 2, 33, device="cuda", dtype=torch.float64))


def test_render_grafx_training_mode_reaches_the_processor_backward_passes():
    """Grad mode through the graph API: in -> ParametricEqualizer -> StereoGain -> out bus with the CUDA processors that
    have a backward pass (grafx_b200/autograd.py).  Forward values equal the in-place (no_grad) render; gradients equal
    the same processors composed by hand; a plan with a forward-only processor raises instead of cutting the graph."""
    import grafx_b200.processors as P
    from grafx_b200.render import mixing_console_plan, render_grafx

    torch.manual_seed(41)
    T, B, L = 3, 2, 6000
    procs = {"eq": P.ParametricEqualizer(num_filters=3, processor_channel="stereo", backend="lfilter").cuda(), "gain": P.StereoGain().cuda()}
    rd = mixing_console_plan(T, ["eq", "gain"])
    x = torch.randn(B, T, 2, L, device="cuda", requires_grad=True)
    prm = {"eq": {k: (0.3 * torch.randn(T, 2, 3, device="cuda")).requires_grad_(True) for k in ("w0", "q_inv", "log_gain")},
           "gain": {"log_gain": (0.3 * torch.randn(T, 2, device="cuda")).requires_grad_(True)}}
    out, inter, buf = render_grafx(procs, x, prm, rd)
    assert out.requires_grad and buf.requires_grad and tuple(buf.shape) == (B, 3 * T + 1, 2, L)
    with torch.no_grad():
        out0, _, buf0 = render_grafx(procs, x, prm, rd)
    assert rel_l2(out.detach().cpu(), out0.cpu()) < 1e-6 and rel_l2(buf.detach().cpu(), buf0.cpu()) < 1e-6
    w = torch.randn_like(out)
    leaves = [x] + list(prm["eq"].values()) + [prm["gain"]["log_gain"]]
    grads = torch.autograd.grad((out * w).sum(), leaves)
    # the same computation by hand (batch-major rows here, node-major inside render_grafx)
    rows = x.reshape(B * T, 2, L)
    rep = lambda t: t.unsqueeze(0).expand(B, *t.shape).reshape(B * T, *t.shape[1:])  # noqa: E731
    y = procs["gain"](procs["eq"](rows, **{k: rep(v) for k, v in prm["eq"].items()}), rep(prm["gain"]["log_gain"]))
    ref = y.reshape(B, T, 2, L).sum(1, keepdim=True)
    ref_grads = torch.autograd.grad((ref * w).sum(), leaves)
    for g, r in zip(grads, ref_grads):
        assert rel_l2(g.cpu(), r.cpu()) < 1e-4, rel_l2(g.cpu(), r.cpu())
    procs2 = {"eq": procs["eq"], "gain": P.Compressor(energy_smoother="ballistics").cuda()}
    prm2 = {"eq": prm["eq"], "gain": {k: torch.zeros(T, v, device="cuda", requires_grad=True) for k, v in procs2["gain"].parameter_size().items()}}
    with pytest.raises(NotImplementedError):
        render_grafx(procs2, x, prm2, rd)


# ------------------------------------------------------------------ training mode of the dynamics processors and the reverb
def _grads_vs_float64(y, leaves, y64, leaves64, w, names, tol):
    g = torch.autograd.grad((y * w.cuda()).sum(), leaves)
    g64 = torch.autograd.grad((y64 * w.double()).sum(), leaves64)
    for name, a, b in zip(names, g, g64):
        r = rel_l2(a.cpu(), b.float())
        assert r <= tol, (name, r)
    return g64


@pytest.mark.parametrize("kind,knee,gain_smoother,in_log", [("compressor", "quadratic", None, False), ("noisegate", "quadratic", None, False),
                                                          ("compressor", "exponential", "iir", True), ("noisegate", "hard", "iir", False)])
def test_dynamics_training_mode_vs_float64_autograd(kind, knee, gain_smoother, in_log):
    """Compressor / NoiseGate with one-pole smoothers in grad mode (grafx_b200/training.py: PyTorch statements + the
    differentiable FIR engine): forward equals the fused no-grad kernel, gradients equal float64 autograd through the
    oracle."""
    from oracle import grafx_oracle as O
    import grafx_b200.processors as P

    gen = torch.Generator().manual_seed(hash((kind, knee)) % 1000)
    B, C, L, N = 3, 2, 5000, 512
    cls = P.Compressor if kind == "compressor" else P.NoiseGate
    proc = cls(knee=knee, gain_smoother=gain_smoother, gain_smooth_in_log=in_log, iir_len=N).cuda()
    x = torch.randn(B, C, L, generator=gen)
    prm = {k: 0.5 * torch.randn(B, v, generator=gen) for k, v in proc.parameter_size().items()}
    xc = x.cuda().requires_grad_(True)
    pc = {k: v.cuda().requires_grad_(True) for k, v in prm.items()}
    y = proc(xc, **pc)
    with torch.no_grad():
        y0 = proc(xc, **pc)
    assert rel_l2(y.detach().cpu(), y0.cpu()) < 2e-5
    x64 = x.double().requires_grad_(True)
    p64 = {k: v.double().requires_grad_(True) for k, v in prm.items()}
    y64 = O.dynamics(kind, x64, **p64, energy_smoother="iir", gain_smoother=gain_smoother, gain_smooth_in_log=in_log, knee=knee, iir_len=N)
    w = torch.randn(B, C, L, generator=gen)
    names = ["x"] + list(prm)
    _grads_vs_float64(y, [xc] + [pc[k] for k in prm], y64, [x64] + [p64[k] for k in prm], w, names, 2e-3)


def test_reverb_training_mode_vs_float64_autograd():
    from oracle import grafx_oracle as O
    import grafx_b200.processors as P

    gen = torch.Generator().manual_seed(77)
    B, L, N = 2, 9000, 6000
    for mode in ("pseudo_midside", "midside"):
        proc = P.STFTMaskedNoiseReverb(ir_len=N, processor_channel=mode).cuda()
        x = torch.randn(B, 2, L, generator=gen)
        prm = {k: 0.5 * torch.randn(B, *v, generator=gen) for k, v in proc.parameter_size().items()}
        xc = x.cuda().requires_grad_(True)
        pc = {k: v.cuda().requires_grad_(True) for k, v in prm.items()}
        y = proc(xc, **pc)
        with torch.no_grad():
            y0 = proc(xc, **pc)
        assert rel_l2(y.detach().cpu(), y0.cpu()) < 2e-5
        x64 = x.double().requires_grad_(True)
        p64 = {k: v.double().requires_grad_(True) for k, v in prm.items()}
        y64 = O.stft_masked_noise_reverb(x64, **p64, ir_len=N, processor_channel=mode)
        w = torch.randn(B, 2, L, generator=gen)
        g64 = _grads_vs_float64(y, [xc] + [pc[k] for k in prm], y64, [x64] + [p64[k] for k in prm], w, ["x"] + list(prm), 1e-3)
        # audio-only gradient: the synthesis kernel makes the response, the convolution carries the graph
        y2 = proc(xc, **{k: v.detach() for k, v in pc.items()})
        gx = torch.autograd.grad((y2 * w.cuda()).sum(), xc)[0]
        assert rel_l2(gx.cpu(), g64[0].float()) < 1e-4


def test_render_eq_compressor_reverb_trains_end_to_end():
    """The BASELINE graph shape (in -> eq -> compressor -> reverb -> out) in grad mode through render_grafx: finite
    gradients for every parameter and the sources, forward equal to the in-place render."""
    import grafx_b200.processors as P
    from grafx_b200.render import mixing_console_plan, render_grafx

    torch.manual_seed(5)
    T, B, L = 2, 2, 8000
    procs = {"eq": P.ParametricEqualizer(num_filters=3, processor_channel="stereo", backend="lfilter").cuda(),
             "compressor": P.Compressor(iir_len=1024).cuda(), "reverb": P.STFTMaskedNoiseReverb(ir_len=6000).cuda()}
    rd = mixing_console_plan(T, ["eq", "compressor", "reverb"])
    x = torch.randn(B, T, 2, L, device="cuda", requires_grad=True)
    prm = {t: {k: (0.3 * torch.randn(T, *((v,) if isinstance(v, int) else v), device="cuda")).requires_grad_(True)
               for k, v in p.parameter_size().items()} for t, p in procs.items()}
    out, _, buf = render_grafx(procs, x, prm, rd)
    with torch.no_grad():
        out0, _, buf0 = render_grafx(procs, x, prm, rd)
    assert rel_l2(out.detach().cpu(), out0.cpu()) < 2e-5 and rel_l2(buf.detach().cpu(), buf0.cpu()) < 2e-5
    leaves = [x] + [v for d in prm.values() for v in d.values()]
    grads = torch.autograd.grad(out.square().mean(), leaves)
    nonzero = 0
    for g in grads:
        assert torch.isfinite(g).all()
        nonzero += int(float(g.abs().max()) > 0)
    # (the knee width only matters for samples inside the knee: its gradient may legitimately vanish)
    assert float(grads[0].abs().max()) > 0 and nonzero >= len(grads) - 1


def test_new_entry_points_empty_batch_and_argument_errors():
    """Empty batches return empty tensors without a launch; the raw ABI rejects bad arguments with error codes."""
    import ctypes
    import grafx_b200.functional as F_
    import grafx_b200.processors as P
    from grafx_b200 import _cabi

    e = torch.empty(0, 2, 100, device="cuda")
    assert F_.fir_filter(e, torch.empty(0, 2, 31, device="cuda")).shape == (0, 2, 100)
    assert F_.envelope(torch.empty(0, 100, device="cuda"), torch.empty(0, 1, device="cuda"), "iir").shape == (0, 100)
    comp = P.Compressor(energy_smoother="ballistics").cuda()
    assert comp(e, **{k: torch.empty(0, v, device="cuda") for k, v in comp.parameter_size().items()}).shape == (0, 2, 100)
    L_ = _cabi.lib()
    x = torch.zeros(1, 1, 64, device="cuda")
    assert L_.gfx_envelope_f32(x.data_ptr(), x.data_ptr(), 1, 1, 64, 3, x.data_ptr(), 0, 0, 16, None, 0, None) == -1   # bad smoother
    assert L_.gfx_envelope_f32(x.data_ptr(), x.data_ptr(), 1, 2, 64, 1, x.data_ptr(), 2, 0, 16, None, 0, None) == -1   # detect=2 needs mono
    assert L_.gfx_envelope_f32(x.data_ptr(), x.data_ptr(), 1, 1, 64, 1, x.data_ptr(), 0, 0, 16, None, 0, None) == -2   # no workspace
    assert L_.gfx_fir_filter_f32(x.data_ptr(), None, x.data_ptr(), 1, 1, 1, 64, 8, None, None, 0, None) == -1
    assert L_.gfx_fir_set_sweep_mb(1) == -1 and L_.gfx_fir_set_mac_form(7) != 0
    assert L_.gfx_dynamics_set_tuning(48) == -1 and L_.gfx_dynamics_set_ballistics_mode(5) == -1
    assert L_.gfx_fma_probe_f32(None, 16, None) == -1


# ------------------------------------------------------------------ first render order reading the caller's sources
@pytest.mark.parametrize("L", [8192 * 3, 5000, 4099, 2050])
@pytest.mark.parametrize("batch", [None, 1, 3])
def test_first_order_source_fold_matches_the_copy_path(batch, L):
    """render_grafx, first render order = a biquad-cascade processor on exactly the source slice: the cascade kernel reads
    the caller's sources and fills the buffer's source slice on the way (gfx_biquad_cascade_ex_f32) instead of a separate
    copy pass.  Output, intermediates and the WHOLE signal buffer must be bit-identical to the copy path (the same kernel
    on the same values), for full tiles, ragged tails and rows that are not 16-byte aligned, 3-D and 4-D sources."""
    import grafx_b200.processors as P
    from grafx_b200 import functional as F_
    from grafx_b200.render import mixing_console_plan, render_grafx

    torch.manual_seed(77)
    T = 5
    shape = (T, 2, L) if batch is None else (batch, T, 2, L)
    x = torch.randn(*shape, device="cuda")
    eq, comp = P.ParametricEqualizer(num_filters=4, processor_channel="stereo", backend="lfilter").cuda(), P.Compressor().cuda()
    prm = {"eq": {k: 0.5 * torch.randn(T, *v, device="cuda") for k, v in eq.parameter_size().items()},
           "compressor": {k: torch.randn(T, v, device="cuda") for k, v in comp.parameter_size().items()}}
    rd = mixing_console_plan(T, ["eq", "compressor"])
    procs = {"eq": eq, "compressor": comp}
    taken = []
    orig = F_.source_fold.__exit__

    def spy(self, *exc):
        taken.append(self.used)
        return orig(self, *exc)

    F_.source_fold.__exit__ = spy
    try:
        out_f, _, buf_f = render_grafx(procs, x, prm, rd)
    finally:
        F_.source_fold.__exit__ = orig
    assert taken == [True]
    eq._gfx_no_source_fold = True   # the copy path: node_copy / copy_ first, plain cascade
    out_c, _, buf_c = render_grafx(procs, x, prm, rd)
    del eq._gfx_no_source_fold
    assert torch.equal(out_f, out_c)
    assert torch.equal(buf_f, buf_c)
    src = buf_f[:, :T] if batch is not None else buf_f[:T]
    assert torch.equal(src, x)


def test_first_order_source_fold_declined_falls_back_to_the_copy():
    """A processor that offers `folds_source_read` but whose first op on the sources is not the cascade kernel (here: a
    mid/side conversion in front of it) leaves the offer unused: render_grafx fills the source slice itself, runs the
    processor again and stops offering to that module."""
    import grafx_b200.processors as P
    from grafx_b200.render import mixing_console_plan, render_grafx
    import torch.nn as nn

    class MidFirst(nn.Module):
        def __init__(self):
            super().__init__()
            self.eq = P.ParametricEqualizer(num_filters=3, processor_channel="stereo", backend="lfilter")

        def folds_source_read(self):
            return True

        def forward(self, input_signals, **kw):
            from grafx_b200 import functional as F_
            return F_.ms_to_lr(self.eq(F_.lr_to_ms(input_signals), **kw))

        def parameter_size(self):
            return self.eq.parameter_size()

    torch.manual_seed(78)
    T, L = 3, 6000
    x = torch.randn(2, T, 2, L, device="cuda")
    proc = MidFirst().cuda()
    prm = {"eq": {k: 0.5 * torch.randn(T, *v, device="cuda") for k, v in proc.parameter_size().items()}}
    rd = mixing_console_plan(T, ["eq"])
    out, _, buf = render_grafx({"eq": proc}, x, prm, rd)
    assert getattr(proc, "_gfx_no_source_fold", False)
    assert torch.equal(buf[:, :T], x)
    ref = P.ParametricEqualizer(num_filters=3, processor_channel="midside", backend="lfilter").cuda()
    out_r, _, buf_r = render_grafx({"eq": ref}, x, prm, rd)
    assert torch.equal(buf, buf_r) and torch.equal(out, out_r)


def test_cascade_ex_abi_rejects_bad_arguments():
    from grafx_b200 import _cabi

    L_ = _cabi.lib()
    x = torch.randn(2, 3, 1, 64, device="cuda")
    y = torch.empty(6, 2, 64, device="cuda")
    c = torch.randn(6, 2, 1, 3, device="cuda")
    ws = torch.empty(1 << 16, dtype=torch.uint8, device="cuda")
    args = lambda xcopy, c_filt, so, si, rep: (x.data_ptr(), xcopy, y.data_ptr(), c.data_ptr(), c.data_ptr(), 6, 1, c_filt, 1, 64,  # noqa: E731
                                               so, si, rep, ws.data_ptr(), ws.numel(), 0)
    assert L_.gfx_biquad_cascade_ex_f32(*args(y.data_ptr(), 2, 2, 3, 1)) != 0   # source form with c_sig (1) < c_filt (2)
    assert L_.gfx_biquad_cascade_ex_f32(*args(None, 1, 2, 3, 1)) != 0           # source mapping without a copy target
    assert L_.gfx_biquad_cascade_ex_f32(*args(y.data_ptr(), 1, 2, 2, 1)) != 0   # src_outer * src_inner != batch
    assert L_.gfx_biquad_cascade_ex_f32(*args(None, 1, 0, 0, 4)) != 0           # batch not a multiple of coef_repeat
    assert L_.gfx_biquad_cascade_ex_f32(*args(None, 1, 0, 0, 0)) != 0
    st = (_cabi.DynamicsStage * 1)()
    assert L_.gfx_dynamics_rep_f32(x.data_ptr(), y.data_ptr(), 6, 1, 64, st, 1, 16, 4, ws.data_ptr(), ws.numel(), 0) != 0  # 6 % 4


@pytest.mark.parametrize("batch,L", [(3, 20000), (2, 4099), (1, 5000)])
def test_render_parameter_rows_by_repeat_match_the_expansion(batch, L):
    """4-D sources: processors that accept it get the per-node parameter rows UN-expanded (`accepts_parameter_repeat`): the
    cascade and dynamics kernels share a coefficient / parameter row over each run of B batch items (gfx_biquad_cascade_ex_f32,
    gfx_dynamics_rep_f32; only their tables kernels know), the reverb synthesises each response once.  Output and the whole
    signal buffer must be bit-identical to the expansion upstream makes (render/graph.py:132-147), here forced by
    switching the offer off on every module."""
    import grafx_b200.processors as P
    from grafx_b200.render import mixing_console_plan, render_grafx

    torch.manual_seed(79 + batch)
    T = 4
    x = torch.randn(batch, T, 2, L, device="cuda")
    procs = {"eq": P.ParametricEqualizer(num_filters=3, processor_channel="midside", backend="lfilter").cuda(),
             "lp": P.LowPassFilter(backend="lfilter").cuda(),
             "compressor": P.Compressor(energy_smoother="ballistics", gain_smoother="iir").cuda(),
             "gate": P.NoiseGate().cuda(),
             "reverb": P.STFTMaskedNoiseReverb(ir_len=20000).cuda()}
    prm = {t: {k: 0.5 * torch.randn(T, *((v,) if isinstance(v, int) else v), device="cuda") for k, v in p.parameter_size().items()}
           for t, p in procs.items()}
    rd = mixing_console_plan(T, ["eq", "lp", "compressor", "gate", "reverb"])
    out_r, _, buf_r = render_grafx(procs, x, prm, rd)
    try:
        for p in procs.values():
            p.accepts_parameter_repeat = lambda: False
        out_e, _, buf_e = render_grafx(procs, x, prm, rd)
    finally:
        for p in procs.values():
            del p.accepts_parameter_repeat
    assert torch.equal(out_r, out_e)
    assert torch.equal(buf_r, buf_e)
