"""CPU: pins the oracle restatement (oracle/grafx_oracle.py) against golden vectors produced by
executing the reference's own code (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from _golden import fixture_names, load, max_rel, oracle_call, oracle_gradients, rel_l2

TOL = 2e-5  # fp32 restatement vs fp32 reference: only op-ordering noise is allowed


@pytest.mark.parametrize("name", fixture_names())
def test_oracle_matches_reference_fixture(name):
    x, params, meta, y_ref, extra = load(name)
    y = oracle_call(name, x, params, meta["kwargs"], extra=extra)
    assert y.shape == y_ref.shape
    assert rel_l2(y, y_ref) < TOL, (name, rel_l2(y, y_ref))
    assert max_rel(y, y_ref) < 10 * TOL


@pytest.mark.parametrize("name", fixture_names(["peq_lfilter", "cfg1_biquad_lfilter", "compressor_iir_None", "reverb_pseudo"]))
def test_oracle_float64_independent_path(name):
    """float64 run through scipy.signal.lfilter / restated istft: an independent second opinion."""
    x, params, meta, y_ref, _ = load(name)
    y = oracle_call(name, x, params, meta["kwargs"], dtype=torch.float64)
    assert rel_l2(y, y_ref) < 1e-4


def test_kat_iir_float64():
    """The reference's only known-answer test (tests/processors/test_filter.py:215-233)."""
    from oracle import grafx_oracle as O
    from _golden import GOLDEN
    import os

    z = np.load(os.path.join(GOLDEN, "kat_iir_f64.npz"))
    x, Bs, As, y = (torch.from_numpy(z[k]) for k in ("x", "Bs", "As", "y"))
    assert torch.allclose(O.iir_lfilter(x, Bs, As, use_torchaudio=False), y, rtol=1e-9, atol=1e-9)


def test_truncated_one_pole_identity():
    from oracle import grafx_oracle as O

    torch.manual_seed(0)
    u = torch.rand(3, 3000, dtype=torch.float64)
    z = torch.tensor([[6.0], [0.3], [3.0]], dtype=torch.float64)
    a = O.truncated_one_pole(u, z, iir_len=256)
    b = O.truncated_one_pole_recursive(u, z, iir_len=256)
    assert (a - b).abs().max() < 1e-12


def test_convolve_guard_and_direct():
    from oracle import grafx_oracle as O

    x, params, meta, y_ref, extra = load("convolve_causal_zerophase")
    h = params["h"]
    assert rel_l2(O.convolve(x, h, "causal"), y_ref) < 1e-6
    assert rel_l2(O.convolve(x, h, "zerophase"), torch.from_numpy(extra["y_zerophase"])) < 1e-6
    assert rel_l2(O.convolve_direct(x, h, "causal").float(), y_ref) < 1e-5
    # the unguarded reference is far away -- documented, not hidden
    assert float(extra["as_shipped_rel_l2"]) > 1e-2


def test_istft_restatement():
    from oracle import grafx_oracle as O

    torch.manual_seed(0)
    p = [0.5 * torch.randn(2, 2, 193) for _ in range(2)]
    a = O.reverb_ir(*p, ir_len=1920, use_torch_istft=True)
    b = O.reverb_ir(*p, ir_len=1920, use_torch_istft=False)
    assert rel_l2(a, b) < 1e-5


@pytest.mark.parametrize("name", fixture_names(["grad_"]))
def test_oracle_gradients_match_reference_autograd(name):
    """Backward fixtures (oracle/make_golden_grad.py: the reference under PyTorch autograd, float32) against the
    float64 gradients of the restatement."""
    x, params, meta, y_ref, extra = load(name)
    w = torch.from_numpy(extra["w"])
    y, gx, gp = oracle_gradients(name, x, params, meta["kwargs"], w)
    assert rel_l2(y, y_ref) < 1e-4
    assert rel_l2(gx, torch.from_numpy(extra["gx"])) < 1e-4
    for k, g in gp.items():
        ref = torch.from_numpy(extra["g_" + k])
        assert max_rel(g, ref) < 1e-3, (name, k, max_rel(g, ref))


# ------------------------------------------------------------------ Ballistics (torchcomp.compressor_core) pins
def _ballistics_inputs(dtype=torch.float64, B=6, L=4000):
    g = torch.Generator().manual_seed(7)
    u = (torch.rand(B, L, generator=g) * 2.5).to(dtype)           # crosses the state (starts at 1) both ways
    z = torch.randn(B, 2, generator=g).to(dtype)
    return u, z


def test_ballistics_oracle_equals_upstream_kernel_text():
    """oracle.ballistics (and the shim the fixtures were generated with) == the torchcomp kernel as restated in
    oracle/torchcomp_core.py, bit for bit in float64."""
    from oracle import grafx_oracle as O, ref_loader as RL, torchcomp_core as TC

    u, z = _ballistics_inputs()
    ts = torch.sigmoid(z)
    zi = torch.ones(u.shape[0], dtype=u.dtype)
    y_kernel = TC.compressor_core(u, zi, ts[:, 0], ts[:, 1])
    assert torch.equal(O.ballistics(u, z), y_kernel)
    assert torch.equal(RL.compressor_core_loop(u, zi, ts[:, 0], ts[:, 1]), y_kernel)
    u32, z32 = _ballistics_inputs(torch.float32)
    ts32 = torch.sigmoid(z32)
    y32 = TC.compressor_core(u32, torch.ones(u32.shape[0]), ts32[:, 0], ts32[:, 1])
    assert rel_l2(O.ballistics(u32, z32), y32) < 1e-6


def test_ballistics_equal_coefficients_is_the_linear_one_pole():
    """Reference-side invariant: with at == rt no branch is involved and the recursion must be the one-pole
    y[t] = (1 - a) y[t-1] + a u[t], y[-1] = 1 (scipy.signal.lfilter with that initial state)."""
    import scipy.signal
    from oracle import grafx_oracle as O

    u, z = _ballistics_inputs()
    z[:, 1] = z[:, 0]
    y = O.ballistics(u, z).numpy()
    a = torch.sigmoid(z[:, 0]).numpy()
    for r in range(u.shape[0]):
        ref, _ = scipy.signal.lfilter([a[r]], [1.0, -(1.0 - a[r])], u[r].numpy(), zi=[(1.0 - a[r]) * 1.0])
        assert np.abs(y[r] - ref).max() < 1e-12


def test_ballistics_single_branch_inputs():
    """An input that stays below the state only ever uses `at` (column 0), one that stays above only `rt`
    (column 1): each branch alone is the same linear one-pole."""
    import scipy.signal
    from oracle import grafx_oracle as O

    _, z = _ballistics_inputs()
    L = 3000
    below = torch.zeros(z.shape[0], L, dtype=torch.float64)           # state decays from 1 towards 0: u < y always
    above = torch.full((z.shape[0], L), 3.0, dtype=torch.float64)     # state rises from 1 towards 3: u > y always
    for u, col in ((below, 0), (above, 1)):
        y = O.ballistics(u, z).numpy()
        c = torch.sigmoid(z[:, col]).numpy()
        for r in range(z.shape[0]):
            ref, _ = scipy.signal.lfilter([c[r]], [1.0, -(1.0 - c[r])], u[r].numpy(), zi=[(1.0 - c[r]) * 1.0])
            assert np.abs(y[r] - ref).max() < 1e-12


def test_envelope_modules_oracle_statements():
    """TruncatedOnePoleIIRFilter / Ballistics / the envelope followers as the oracle states them are the reference's
    own modules (core/envelope.py, dynamics.py:745-790) on the same inputs -- golden file made by
    oracle/make_golden_envelope.py."""
    import os
    from _golden import GOLDEN
    from oracle import grafx_oracle as O

    zf = np.load(os.path.join(GOLDEN, "envelope_modules.npz"))
    u, x = torch.from_numpy(zf["u"]), torch.from_numpy(zf["x"])
    z1, z2 = torch.from_numpy(zf["z1"]), torch.from_numpy(zf["z2"])
    assert rel_l2(O.truncated_one_pole(u, z1, iir_len=int(zf["iir_len"])), torch.from_numpy(zf["y_onepole"])) < TOL
    assert rel_l2(O.ballistics(u, z2), torch.from_numpy(zf["y_ballistics"])) < TOL
    for det in ("energy", "amplitude"):
        assert rel_l2(O.envelope_follower(x, z1, "iir", det, iir_len=int(zf["iir_len"])), torch.from_numpy(zf[f"env_iir_{det}"])) < TOL
        assert rel_l2(O.envelope_follower(x, z2, "ballistics", det), torch.from_numpy(zf[f"env_ballistics_{det}"])) < TOL
