"""CPU: host-side logic around the kernels that needs no GPU -- the differentiable statement of the coefficient
design used in training mode, argument validation of the newer C-ABI entry points, workspace sizing of the
reverb geometries, the capture wrapper's refusal of CPU tensors."""
import pytest
import torch


def test_design_statement_dispatch_matches_design_module_and_carries_grad():
    from grafx_b200 import functional as F_
    from grafx_b200.processors import design as D

    torch.manual_seed(0)
    w0, q, g = (torch.randn(3, 2, 4, requires_grad=True) for _ in range(3))
    for shelf in (True, False):
        for a, b in zip(F_.biquad_design("peq", w0, q, g, flags=int(shelf)), D.parametric_eq(w0, q, g, shelf)):
            assert torch.equal(a, b) and a.requires_grad
    for kind in ("peaking", "lowshelf", "highshelf"):
        for a, b in zip(F_.biquad_design(kind, w0, q, g), D.eq_band(kind, w0, q, g)):
            assert torch.equal(a, b)
    for kind in ("lowpass", "highpass", "bandpass", "bandreject", "allpass"):
        for a, b in zip(F_.biquad_design(kind, w0, q), D.simple_filter(kind, w0, q)):
            assert torch.equal(a, b)
    Bs = torch.randn(3, 4, 3, requires_grad=True)
    a1, a2, a0 = (torch.randn(3, 4, requires_grad=True) for _ in range(3))
    for a, b in zip(F_.biquad_design("stable", Bs, a1, a2, a0, flags=2), D.stable_biquad(Bs, a1, a2, a0, True)):
        assert torch.equal(a, b)
    for a, b in zip(F_.biquad_design("stable", Bs, a1, a2, None, flags=0), D.stable_biquad(Bs, a1, a2, None, False)):
        assert torch.equal(a, b)
    five = [torch.randn(3, 4, requires_grad=True) for _ in range(5)]
    num, den = F_.biquad_design("svf", *five)
    (num.sum() + den.sum()).backward()
    assert all(t.grad is not None for t in five)


def test_wants_grad_follows_autograd_state():
    from grafx_b200 import functional as F_

    t = torch.zeros(2, requires_grad=True)
    assert F_._wants_grad(None, t)
    assert not F_._wants_grad(t.detach(), None)
    with torch.no_grad():
        assert not F_._wants_grad(t)
    with pytest.raises(NotImplementedError):
        F_._no_backward("some_op", t)


def test_new_entry_points_validate_arguments_without_gpu():
    from grafx_b200 import _cabi

    L = _cabi.lib()
    assert L.gfx_node_copy_f32(None, None, 1, 1, 4, 4, 4, 4, 4, None) == -1
    assert L.gfx_lag_dots_f32(None, None, None, None, None, 1, 4, 0, None) == -1
    assert L.gfx_row_mean_square_f32(None, None, 1, 4, None) == -1
    assert L.gfx_pointwise_f32(9, None, None, 1, 1, 4, None, None, None, None, None, 0, 0, None) == -1
    # reverb workspace: tuned geometry, general power-of-two geometry, unsupported geometry
    fast = L.gfx_reverb_ir_workspace_bytes(4, 384, 192, 96000)
    general = L.gfx_reverb_ir_workspace_bytes(4, 512, 128, 96000)
    assert 0 < fast < general
    assert general >= 4 * 2 * (1 + 96000 // 128) * 512 * 4
    assert L.gfx_reverb_ir_workspace_bytes(4, 300, 100, 96000) == 0
    assert L.gfx_reverb_ir_workspace_bytes(4, 8192, 2048, 96000) == 0


def test_captured_render_refuses_cpu_tensors():
    from grafx_b200.render import CapturedRender, mixing_console_plan

    with pytest.raises(RuntimeError):
        CapturedRender({}, torch.zeros(1, 2, 2, 64), {}, mixing_console_plan(2, ["eq"]))


def test_bench_gpu_affinity_helper_degrades_without_nvml():
    import bench

    cpus = bench.gpu_local_cpus(0)
    assert cpus is None or len(cpus) > 0


def test_render_grafx_training_mode_is_functional_and_differentiable():
    """Grad mode: the plan is evaluated without in-place writes (host logic checked here with plain torch modules on
    the CPU: slice / sum / scatter / index paths, 3-D and 4-D sources, gradients to sources and parameters)."""
    import torch.nn as nn
    from grafx_b200.render import mixing_console_plan, render_grafx
    from grafx_b200.render.plan import RenderData, _AggregationData as Agg, _SingleRenderData as It, _TensorAccessData as Acc

    class Gain(nn.Module):
        def forward(self, x, g):
            return x * g.view(-1, 1, 1)

    rd = mixing_console_plan(3, ["gain", "gain"])
    for shape in ((3, 2, 16), (4, 3, 2, 16)):
        x = torch.randn(*shape, requires_grad=True)
        g = torch.randn(3, 1, requires_grad=True)
        out, inter, buf = render_grafx({"gain": Gain()}, x, {"gain": {"g": g}}, rd)
        node_axis = 0 if len(shape) == 3 else 1
        ref = (x * (g * g).view(-1, 1, 1)).sum(node_axis, keepdim=True)
        assert torch.allclose(out, ref, atol=1e-6) and inter == []
        assert buf.shape[node_axis] == rd.num_nodes and torch.allclose(buf.narrow(node_axis, 0, 3), x)
        assert torch.allclose(buf.narrow(node_axis, 3, 3), x * g.view(-1, 1, 1), atol=1e-6)
        gx, gg = torch.autograd.grad(out.square().sum(), (x, g))
        rx, rg = torch.autograd.grad(ref.square().sum(), (x, g))
        assert torch.allclose(gx, rx, atol=1e-5) and torch.allclose(gg, rg, atol=1e-4)
    # irregular plan: index read, scatter into two buses, index write
    idx = torch.tensor([2, 0, 3, 1])
    rd2 = RenderData("beam", 10, 2, True, [
        It("in", [Acc("none", ())], [Agg("none")], Acc("slice", (0, 4)), Acc("slice", (0, 4))),
        It("gain", [Acc("index", idx)], [Agg("none")], Acc("index", idx), Acc("index", torch.tensor([4, 5, 6, 7]))),
        It("mix", [Acc("slice", (4, 8))], [Agg("scatter", torch.tensor([0, 1, 1, 0]))], Acc("slice", (0, 2)), Acc("slice", (8, 10)))])
    x = torch.randn(4, 2, 8, requires_grad=True)
    g = torch.randn(4, 1, requires_grad=True)
    out, _, buf = render_grafx({"gain": Gain()}, x, {"gain": {"g": g}}, rd2)
    y = x[idx] * g[idx].view(-1, 1, 1)
    ref = torch.stack([y[0] + y[3], y[1] + y[2]])
    assert torch.allclose(out, ref, atol=1e-6) and torch.allclose(buf[8:10], ref, atol=1e-6)
    assert torch.allclose(torch.autograd.grad(out.sum(), g)[0], torch.autograd.grad(ref.sum(), g)[0], atol=1e-5)


def test_state_dict_keys_match_the_reference():
    """Every drop-in module registers the buffers its upstream class registers (same names, shapes, dtypes): an upstream
    checkpoint loads with strict=True.  Fixture: oracle/make_state_dict_keys.py run against the reference."""
    import json
    import os

    import grafx_b200.processors as P

    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "state_dict_keys.json")))
    checked = 0
    for name, keys in ref.items():
        if "error" in keys:
            continue
        mod = getattr(P, name)()
        ours = {k: [list(v.shape), str(v.dtype)] for k, v in mod.state_dict().items()}
        assert ours == keys, (name, sorted(set(ours) ^ set(keys)))
        checked += 1
    assert checked >= 25


def test_bench_config_is_identical_in_both_arms_and_the_shard_covers_the_batch():
    """bench.py: the `config` object is built by one function for both arms (the driver compares them), and the
    config-5 shard hands every one of the 128 renders to exactly one rank, in whole chunks, for 1 / 2 / 4 / 8 ranks."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for world in (1, 2, 4, 8):
        assert bench.config_of("cfg5", world) == bench.config_of("cfg5", world)
        total = 0
        for rank in range(world):
            g = bench.GraphWorkload(world, rank)
            assert g.n_chunks * g.chunk == g.B_local and g.chunk <= 32
            total += g.B_local
        assert total == 128
        assert bench.config_of("cfg5", world)["renders_per_gpu"] == 128 // world
    for name in ("cfg1", "cfg2", "cfg2lf", "cfg3", "cfg3b", "cfg4", "cfg4b"):
        wl = bench.Workload(name)
        x, prm = wl.host_inputs(B=2)
        assert x.shape == (2, wl.C, wl.L) and wl.alg_bytes() >= 8 * wl.samples()
        assert set(bench.config_of(name, 1)) == {"workload", "batch", "channels", "length", "l2", "parallelism"}


def test_library_staleness_is_decided_by_content_not_mtime():
    """grafx_b200.build.needs_rebuild compares a digest of the sources with the one recorded at build time: a copy of the
    tree that does not preserve timestamps (the snapshot sent to a GPU box) must not trigger a rebuild."""
    import os

    from grafx_b200 import _cabi, build

    _cabi.lib()  # (builds if needed)
    assert os.path.exists(build.LIB + ".srchash") and not build.needs_rebuild()
    src = os.path.join(build.CSRC, "abi.cu")
    st = os.stat(src)
    try:
        os.utime(src, (st.st_atime, st.st_mtime + 10_000))
        assert not build.needs_rebuild()
    finally:
        os.utime(src, (st.st_atime, st.st_mtime))
