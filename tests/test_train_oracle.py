"""CPU: pin the training-step oracle (oracle/mds_train_oracle.py) against the committed goldens, which
oracle/make_train_golden.py wrote after checking two optimizer steps against the UNMODIFIED reference module
(train mode), its FocalLoss and torch.optim.SGD — and against the reference itself when /root/reference exists."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import mds_oracle as O
from oracle import mds_train_oracle as TO

GOLD = Path(__file__).parent / "golden"
REF = Path("/root/reference")
CASES = [("t3_small", O.ModelConfig(num_frames=9), 2, (3, 5)), ("t11_small", O.ModelConfig(num_frames=33), 2, (4, 6))]


@pytest.mark.parametrize("tag,cfg,b,hw", CASES)
def test_train_oracle_matches_golden(tag, cfg, b, hw):
    g = np.load(GOLD / f"train_{tag}.npz")
    sd = O.make_state_dict(cfg, seed=1234, calib_hw=(96, 160))
    enc, targets = TO.make_case(cfg, b, hw, seed=7)
    loss, logits, grads, stats = TO.loss_and_grads(sd, enc, targets, cfg, torch.from_numpy(g["dp"]), torch.from_numpy(g["do"]), 0.4, 1.2)
    np.testing.assert_allclose(loss.numpy(), g["loss"], rtol=1e-4)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-4, atol=1e-4 * np.abs(g["logits"]).max())
    assert set(TO.trainable_keys(cfg)) == set(grads)
    for k, v in grads.items():
        n = float(g["gradnorm:" + k])
        assert abs(float(v.double().norm()) - n) <= 1e-3 * n + 1e-7, k
        np.testing.assert_allclose(v.flatten()[:16].numpy(), g["gradhead:" + k], rtol=0, atol=2e-3 * float(v.abs().max()) + 1e-8)
    for k, v in stats.items():
        np.testing.assert_allclose(v.flatten()[:16].numpy(), g["stathead:" + k], rtol=1e-4, atol=1e-6)


def test_train_pin_report_matches_reference():
    rep = json.loads((GOLD / "train_pin_report.json").read_text())
    assert len(rep) == 4
    for k, v in rep.items():
        assert v["max_rel_err"] < 2e-4, k        # loss, logits, every gradient, BN statistic and SGD-updated parameter


def test_sgd_and_scaler_semantics():
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(50, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([ref], lr=0.1, momentum=0.9, nesterov=True)
    params, bufs = {"w": p0.clone()}, {}
    for _ in range(4):
        grad = torch.randn(50, generator=g)
        ref.grad = grad.clone()
        opt.step()
        TO.sgd_nesterov_step(params, {"w": grad}, bufs, 0.1, 0.9)
    assert torch.allclose(params["w"], ref.detach(), atol=1e-6)
    s = TO.GradScalerOracle(init_scale=1024.0, growth_interval=2)
    s.update(True); assert s.scale == 512.0 and s.tracker == 0
    s.update(False); s.update(False); assert s.scale == 1024.0 and s.tracker == 0


def test_focal_loss_gradient_closed_form():
    """The closed form used by head_train_kernel (csrc/train.cuh) equals autograd of src/losses.py's formula."""
    x = torch.tensor([[-2.0, 0.3], [1.5, -0.1], [0.0, 4.0]], requires_grad=True)
    t = torch.tensor([[0.0, 1.0], [0.7, 0.0], [1.0, 0.2]])
    a, gm = 0.4, 1.2
    TO.sigmoid_focal_loss(x, t, a, gm).backward()
    p = torch.sigmoid(x.detach())
    ce = torch.nn.functional.binary_cross_entropy_with_logits(x.detach(), t, reduction="none")
    q = p + t - 2 * p * t
    at = a * t + (1 - a) * (1 - t)
    closed = at * ((p - t) * q ** gm + ce * gm * q ** (gm - 1) * (1 - 2 * t) * p * (1 - p)) / x.numel()
    assert torch.allclose(x.grad, closed, atol=1e-7)


@pytest.mark.skipif(not REF.exists(), reason="reference tree only exists in the authoring container")
def test_train_oracle_vs_unmodified_reference_step():
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
    import make_train_golden as MT
    cfg = O.ModelConfig(num_frames=9)
    sd = O.make_state_dict(cfg, seed=5, calib_hw=(96, 160))
    enc, targets = TO.make_case(cfg, 2, (3, 4), seed=9)
    (r_loss, r_logits, r_grads, _), = MT.reference_steps(cfg, sd, enc, targets, [21])
    dp, do = MT.replay_masks(cfg, 2, 21)
    loss, logits, grads, _ = TO.loss_and_grads(sd, enc, targets, cfg, dp, do, MT.ALPHA, MT.GAMMA)
    assert torch.allclose(loss, r_loss, rtol=1e-5) and torch.allclose(logits, r_logits, rtol=1e-5, atol=1e-6)
    for k in grads:
        assert torch.allclose(grads[k], r_grads[k], rtol=1e-4, atol=1e-7), k


def test_batchnorm_silu_backward_closed_form():
    """The algebra of bn_bwd_reduce / bn_bwd_finalize / bn_bwd_apply (csrc/train.cuh): with dz = da * silu'(z),
    S1 = sum(dz), S2 = sum(dz * (y - mean)):  d beta = S1, d gamma = rstd * S2, and
    dy = gr * dz + A * y + B  with gr = gamma * rstd, A = -rstd * gr * (d gamma / M), B = -gr * S1 / M - mean * A."""
    g = torch.Generator().manual_seed(0)
    M, C = 257, 24
    y = (torch.randn(M, C, generator=g) * 1.3 + 0.4).double().requires_grad_(True)
    gamma = (torch.rand(C, generator=g) + 0.5).double().requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.1).double().requires_grad_(True)
    da = torch.randn(M, C, generator=g).double()
    eps = 1e-5
    mean, var = y.mean(0), y.var(0, unbiased=False)
    rstd = 1.0 / torch.sqrt(var + eps)
    z = (y - mean) * rstd * gamma + beta
    (torch.nn.functional.silu(z) * da).sum().backward()
    with torch.no_grad():
        sig = torch.sigmoid(z)
        dz = da * sig * (1 + z * (1 - sig))
        S1, S2 = dz.sum(0), (dz * (y - mean)).sum(0)
        dgamma = rstd * S2
        gr = gamma * rstd
        A = -rstd * gr * (dgamma / M)
        B = -gr * S1 / M - mean * A
        dy = gr * dz + A * y + B
    assert torch.allclose(beta.grad, S1, rtol=1e-10, atol=1e-12)
    assert torch.allclose(gamma.grad, dgamma, rtol=1e-10, atol=1e-12)
    assert torch.allclose(y.grad, dy, rtol=1e-9, atol=1e-12)


def test_gem_backward_closed_form():
    """GeM with learnable p (multidim_stacker.py:42-45): f = m^(1/p), m = mean(c^p), c = max(x, eps);
    df/dx_i = m^(1/p - 1) c_i^(p-1) / P  (x_i >= eps),  df/dp = f * (-ln m / p^2 + mean(c^p ln c) / (p m))
    -- the coefficients head_grad_kernel / gem_bwd_kernel use."""
    g = torch.Generator().manual_seed(1)
    P = 37
    x = (torch.rand(P, generator=g) * 2 - 0.3).double().requires_grad_(True)
    p = torch.tensor(3.0, dtype=torch.double, requires_grad=True)
    eps = 1e-6
    f = x.clamp(min=eps).pow(p).mean().pow(1.0 / p)
    f.backward()
    with torch.no_grad():
        c = x.clamp(min=eps)
        m = c.pow(p).mean()
        mlog = (c.pow(p) * c.log()).mean()
        dx = torch.where(x >= eps, m.pow(1.0 / p - 1) * c.pow(p - 1) / P, torch.zeros_like(x))
        dp = f * (-m.log() / p ** 2 + mlog / (p * m))
    assert torch.allclose(x.grad, dx, rtol=1e-10, atol=1e-14)
    assert torch.allclose(p.grad, dp, rtol=1e-10)
