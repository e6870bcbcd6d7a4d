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
