"""GPU: the CUDA training step (mds_train_step, BASELINE.json configs[4]) against the CPU fp32 oracle on identical
inputs, weights and DropPath / Dropout masks.

Tolerances.  The step stores activations and activation-gradients in fp16 with fp32 accumulation (the reference trains
under fp16 autocast + GradScaler, src/argus_models.py:35-36,56-60), so a gradient tensor carries ~1e-3..1e-2 relative
noise: GRAD_TOL bounds the relative L2 error of every gradient tensor, LOSS_TOL the loss, LOGIT_TOL the logits,
STAT_TOL the BatchNorm running statistics (fp32 sums of fp16 conv outputs)."""
import pytest
import torch

from oracle import mds_oracle as O
from oracle import mds_train_oracle as TO
import train_parity as TP

pytestmark = pytest.mark.gpu
GRAD_TOL, LOSS_TOL, LOGIT_TOL, STAT_TOL = 2e-2, 2e-3, 5e-3, 2e-3
DEV = "cuda:0"


def _check(errs):
    bad = {}
    for k, v in errs.items():
        tol = LOSS_TOL if k == "loss" else LOGIT_TOL if k == "logits" else STAT_TOL if k.startswith("stat:") else GRAD_TOL
        if not v <= tol:
            bad[k] = (v, tol)
    assert not bad, bad


@pytest.mark.parametrize("cfg,b,hw", [(O.ModelConfig(num_frames=9), 2, (3, 5)),
                                      (O.ModelConfig(num_frames=33), 2, (4, 6)),
                                      (O.ModelConfig(num_frames=15), 3, (23, 40))])
def test_forward_backward_matches_oracle(cfg, b, hw):
    _check(TP.compare(cfg, b, hw))


def test_wide_feature_map_uses_column_tiles_and_row_chunks():
    """43 columns -> two 40-column tiles in the depthwise kernels; 13 rows -> several row chunks in the weight gradient."""
    _check(TP.compare(O.ModelConfig(num_frames=9), 2, (13, 43)))


def test_forward_backward_without_drops():
    _check(TP.compare(O.ModelConfig(num_frames=9), 3, (5, 7), drop=False))


def test_long_sequence_config_full_feature_map():
    """configs[4]: 33 frames (T = 11), batch 4, 23 x 40 feature map of a 1280 x 736 frame."""
    _check(TP.compare(O.ModelConfig(num_frames=33), 4, (23, 40)))


def test_three_sgd_steps_follow_the_oracle():
    cfg, b, hw, lr = O.ModelConfig(num_frames=9), 2, (4, 5), 0.05
    sd = O.make_state_dict(cfg, seed=1234, calib_hw=(96, 160))
    net, tr = TP.build(cfg, sd, lr=lr)
    cur = {k: v.clone() for k, v in sd.items()}
    bufs = {}
    for step in range(3):
        enc, targets = TO.make_case(cfg, b, hw, seed=100 + step)
        enc = enc.half().float()
        dp, do = TO.make_masks(cfg, b, 0.2, 0.2, 200 + step)
        _, _, grads, stats = TO.loss_and_grads(cur, enc, targets, cfg, dp, do, TP.ALPHA, TP.GAMMA)
        params = {k: cur[k] for k in grads}
        TO.sgd_nesterov_step(params, grads, bufs, lr)
        cur.update(params)
        cur.update(stats)
        tr.step_on_features(TP.to_nhwc16(enc, b, cfg.num_stacks), targets, dp, do, apply_update=True)
    scale, tracker, found_inf, steps = tr.scaler_state()
    assert steps == 3 and found_inf == 0 and scale == 65536.0 and tracker == 3
    worst = {}
    for k in TO.trainable_keys(cfg):
        delta_ref = cur[k].flatten() - sd[k].flatten()
        delta = tr.get(k) - sd[k].flatten()
        floor = 1e-3 * lr * delta_ref.numel() ** 0.5
        worst[k] = TP.rel_l2(delta, delta_ref, floor)
    bad = {k: v for k, v in worst.items() if v > 5e-2}
    assert not bad, bad
    for k in tr.buffer_names:
        assert TP.rel_l2(tr.get(k), cur[k]) < STAT_TOL, k
    # write-back into the nn.Module: state_dict carries the trained values and the BatchNorm step counter
    tr.sync_to_module()
    msd = net.state_dict()
    assert int(msd["conv3d_encoder.0.bn1.bn3d.num_batches_tracked"]) == 3
    assert torch.equal(msd["classifier.weight"].cpu().flatten(), tr.get("classifier.weight"))
    tr.close()


def test_step_is_bit_reproducible():
    cfg, b, hw = O.ModelConfig(num_frames=9), 2, (6, 8)
    sd = O.make_state_dict(cfg, seed=1234, calib_hw=(96, 160))
    enc, targets = TO.make_case(cfg, b, hw, seed=3)
    dp, do = TO.make_masks(cfg, b, 0.2, 0.2, 4)
    outs = []
    for _ in range(2):
        net, tr = TP.build(cfg, sd)
        loss, _ = tr.step_on_features(TP.to_nhwc16(enc, b, cfg.num_stacks), targets, dp, do, apply_update=False)
        outs.append((loss.item(), [tr.get(k, "grad") for k in tr.param_names]))
        tr.close()
    assert outs[0][0] == outs[1][0]
    for a, c in zip(outs[0][1], outs[1][1]):
        assert torch.equal(a, c)


def test_grad_scaler_skips_the_step_on_overflow():
    """GradScaler semantics (argus_models.py:65-66): a non-finite gradient skips the update and halves the scale."""
    cfg, b, hw = O.ModelConfig(num_frames=9), 2, (3, 5)
    sd = O.make_state_dict(cfg, seed=1234, calib_hw=(96, 160))
    net, tr = TP.build(cfg, sd, init_scale=2.0 ** 60)       # fp16 gradient tensors overflow
    enc, targets = TO.make_case(cfg, b, hw, seed=3)
    before = {k: tr.get(k) for k in tr.param_names}
    tr.step_on_features(TP.to_nhwc16(enc, b, cfg.num_stacks), targets, apply_update=True)
    scale, tracker, found_inf, steps = tr.scaler_state()
    assert steps == 0 and scale == 2.0 ** 59 and tracker == 0 and found_inf == 0
    for k, v in before.items():
        assert torch.equal(tr.get(k), v), k
    tr.close()


def test_internal_masks_and_train_step_api():
    """Masks drawn inside the library (no mask tensors passed) and the reference-shaped train_step(batch, state)."""
    cfg = O.ModelConfig(num_frames=9)
    sd = O.make_state_dict(cfg, seed=1234, calib_hw=(96, 160))
    net, tr = TP.build(cfg, sd, lr=0.01)
    frames = torch.randint(0, 256, (2, 9, 96, 160), dtype=torch.uint8, generator=torch.Generator().manual_seed(0))
    target = torch.tensor([[1.0, 0.0], [0.0, 0.5]])
    losses = [tr.train_step((frames, target))["loss"] for _ in range(8)]
    out = tr.train_step((frames, target))
    assert out["prediction"].shape == (2, 2) and out["target"].shape == (2, 2)
    assert all(l == l and l < 10 for l in losses)
    assert min(losses[4:]) < losses[0]                      # the same batch repeated: the loss goes down
    # the encoder features used by the step equal the inference path's encoder output
    f = tr.encoder_features(frames.to(DEV))
    assert f.shape == (2, 3, 3, 5, 192) and torch.isfinite(f.float()).all()
    tr.close()


def test_model_ema_follows_the_reference_recursion():
    """ModelEma.update (src/ema.py:49-57): ema = decay * ema + (1 - decay) * value in float32, seeded with the initial model;
    bit-exact against the same recursion evaluated by torch on the parameter values read back after every step."""
    cfg, b, hw, decay = O.ModelConfig(num_frames=9), 2, (4, 5), 0.9
    sd = O.make_state_dict(cfg, seed=1234, calib_hw=(96, 160))
    from ball_action_spotting_b200 import FrozenEncoderTrainer
    net, tr0 = TP.build(cfg, sd)
    tr0.close()
    tr = FrozenEncoderTrainer(net, lr=0.05, ema_decay=decay)
    names = ["conv3d_encoder.1.conv_pw.weight", "classifier.bias", "global_pool.p", "conv3d_encoder.0.bn2.bn3d.running_var"]
    ema = {k: tr.get(k).clone() for k in names}
    for k in names:
        assert torch.equal(tr.get(k, "ema"), ema[k])              # seeded with a copy of the model
    for step in range(3):
        enc, targets = TO.make_case(cfg, b, hw, seed=300 + step)
        tr.step_on_features(TP.to_nhwc16(enc, b, cfg.num_stacks), targets)
        tr.ema_update()
        for k in names:
            ema[k] = decay * ema[k] + (1.0 - decay) * tr.get(k)
            assert torch.equal(tr.get(k, "ema"), ema[k]), (step, k)
    esd = tr.ema_state_dict()
    assert set(esd) == set(net.state_dict())
    assert torch.equal(esd["classifier.bias"].cpu(), ema["classifier.bias"])
    assert int(esd["conv3d_encoder.0.bn1.bn3d.num_batches_tracked"]) == int(0.9 * int(0.9 * int(0.9 * 0 + 0.1 * 1) + 0.1 * 2) + 0.1 * 3)
    tr.close()


def test_val_step_uses_ema_weights_and_matches_the_oracle_loss():
    """BallActionModel.val_step (argus_models.py:76-91): eval forward of the EMA model, focal loss, sigmoid."""
    from ball_action_spotting_b200 import FrozenEncoderTrainer
    cfg = O.ModelConfig(num_frames=9)
    sd = O.make_state_dict(cfg, seed=1234, calib_hw=(96, 160))
    net, tr0 = TP.build(cfg, sd)
    tr0.close()
    tr = FrozenEncoderTrainer(net, lr=0.05, ema_decay=0.5)
    frames = torch.randint(0, 256, (2, 9, 96, 160), dtype=torch.uint8, generator=torch.Generator().manual_seed(0))
    target = torch.tensor([[1.0, 0.0], [0.3, 1.0]])
    for _ in range(2):
        tr.train_step((frames, target))
    out = tr.val_step((frames, target))
    # oracle: eval-mode forward with the EMA state dict, the reference's focal loss, sigmoid
    esd = {k: v.detach().float().cpu() for k, v in tr.ema_state_dict().items()}
    with torch.no_grad():
        ref_logits = O.forward(esd, O.pad_normalize(frames, (160, 96)), cfg)
    ref_loss = TO.sigmoid_focal_loss(ref_logits, target, TP.ALPHA, TP.GAMMA)
    tol = 1e-3 * (920.0 / 15.0) ** 0.5                      # 96x160 input: 15 GeM positions (see tests/test_e2e_gpu.py)
    assert (out["prediction"].cpu() - torch.sigmoid(ref_logits)).abs().max() <= tol
    assert abs(out["loss"] - ref_loss.item()) <= tol * max(1.0, abs(ref_loss.item()))
    assert out["target"].shape == (2, 2)
    # the EMA weights differ from the current ones, and val_step evaluates the EMA ones
    tr.sync_to_module()
    cur = net(frames.to(DEV)).cpu()
    assert (torch.sigmoid(cur) - out["prediction"].cpu()).abs().max() > 1e-6
    tr.close()
