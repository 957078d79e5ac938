"""Pin ``oracle/mds_train_oracle.py`` against the UNMODIFIED reference and write ``tests/golden/train_*.npz``.

Runs only in the authoring container (reads /root/reference).  What is checked, on identical weights / inputs:

* the reference's own ``conv2d_projection`` / ``forward_3d`` / ``forward_head`` (src/models/multidim_stacker.py,
  imported by file path on top of oracle/timm_shim) in ``.train()`` mode, its own ``FocalLoss`` (src/losses.py) and
  ``torch.optim.SGD(momentum=0.9, nesterov=True)`` (configs/ball_action/ball_finetune_long_004.py:51-55) for two
  consecutive steps, against the functional restatement: loss, logits, every gradient, every BN running statistic
  and every updated parameter must agree to float32 round-off;
* the DropPath / Dropout draws: the reference draws them from the global CPU generator; the same seed replayed
  through the same calls (``bernoulli_`` on (b,1,1,1,1), ``F.dropout`` on (b, F)) yields the masks handed to the
  restatement.

Usage: python oracle/make_train_golden.py
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import mds_oracle as O            # noqa: E402
from oracle import mds_train_oracle as TO     # noqa: E402
from oracle.make_golden import REF, load_by_path, load_reference_model_module, nn_module_params  # noqa: E402

GOLD = ROOT / "tests" / "golden"
ALPHA, GAMMA, LR, MOM = 0.4, 1.2, 0.01, 0.9
DP, DO = 0.2, 0.2


def replay_masks(cfg, b, seed):
    torch.manual_seed(seed)
    keep = 1 - DP
    dp = []
    for _ in range(cfg.num_3d_blocks):
        m = torch.empty((b, 1, 1, 1, 1)).bernoulli_(keep)
        m.div_(keep)
        dp.append(m.view(b))
    do = F.dropout(torch.ones(b, cfg.num_features), p=DO, training=True)
    return torch.stack(dp), do


def reference_steps(cfg, sd, enc_feats, targets, seeds):
    ref = load_reference_model_module()
    losses = load_by_path("ref_losses", REF / "src/losses.py")
    m = ref.MultiDimStacker(**nn_module_params(cfg))
    m.load_state_dict(sd, strict=True)
    m.train()
    for p in m.conv2d_encoder.parameters():          # argus_models.py:104-110
        p.requires_grad_(False)
    params = [p for p in m.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=LR, momentum=MOM, nesterov=True)
    crit = losses.FocalLoss(alpha=ALPHA, gamma=GAMMA, reduction="mean")
    T = cfg.num_stacks
    b = enc_feats.shape[0] // T
    out = []
    for seed in seeds:
        torch.manual_seed(seed)
        opt.zero_grad()
        x = m.conv2d_projection(enc_feats).contiguous()                       # multidim_stacker.py:216
        x = x.view(b, T, cfg.num_3d_features, x.shape[-2], x.shape[-1])        # :218
        logits = m.forward_head(m.forward_3d(x))                              # :241-242
        loss = crit(logits, targets)
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.requires_grad}
        opt.step()
        state = {k: v.detach().clone() for k, v in m.state_dict().items()}
        out.append((loss.detach().clone(), logits.detach().clone(), grads, state))
    return out


def main():
    report = {}
    for tag, cfg, b, hw in (("t3_small", O.ModelConfig(num_frames=9), 2, (3, 5)),
                            ("t11_small", O.ModelConfig(num_frames=33), 2, (4, 6))):
        sd = O.make_state_dict(cfg, seed=1234, calib_hw=(96, 160))
        enc, targets = TO.make_case(cfg, b, hw, seed=7)
        seeds = [11, 12]
        ref_steps = reference_steps(cfg, sd, enc, targets, seeds)

        cur = {k: v.clone() for k, v in sd.items()}
        bufs = {}
        worst = 0.0
        gold = {}
        for step, seed in enumerate(seeds):
            dp, do = replay_masks(cfg, b, seed)
            loss, logits, grads, stats = TO.loss_and_grads(cur, enc, targets, cfg, dp, do, ALPHA, GAMMA)
            r_loss, r_logits, r_grads, r_state = ref_steps[step]

            def rel(a, r):
                # a BN bias followed (through convs) by another train-mode BN has an exactly-zero gradient unless a
                # sample was dropped; its computed value is round-off noise (~1e-10), hence the absolute floor
                return ((a - r).abs().max() / r.abs().max().clamp_min(1e-4)).item()
            errs = {"loss": rel(loss, r_loss), "logits": rel(logits, r_logits)}
            assert set(grads) == set(r_grads), (set(grads) ^ set(r_grads))
            for k in grads:
                errs["grad:" + k] = rel(grads[k], r_grads[k])
            params = {k: cur[k] for k in grads}
            TO.sgd_nesterov_step(params, grads, bufs, LR, MOM)
            cur.update(params)
            cur.update(stats)
            for k in r_state:
                if k.startswith("conv2d_encoder."):
                    continue
                if k.endswith("num_batches_tracked"):
                    assert int(cur[k]) == int(r_state[k]), k
                    continue
                errs["state:" + k] = rel(cur[k], r_state[k])
            w = max(errs.values())
            worst = max(worst, w)
            report[f"{tag}.step{step}"] = {"max_rel_err": w, "worst": max(errs, key=errs.get), "loss": float(loss)}
            if step == 0:
                # inputs are regenerated from seeds (TO.make_case); only digests of the outputs are stored
                gold = {"dp": dp.numpy(), "do": do.numpy(), "loss": loss.numpy(), "logits": logits.numpy()}
                for k, v in grads.items():
                    gold["gradnorm:" + k] = v.double().norm().numpy()
                    gold["gradhead:" + k] = v.flatten()[:16].numpy()
                for k, v in stats.items():
                    gold["stathead:" + k] = v.flatten()[:16].numpy()
        assert worst < 2e-4, report
        np.savez_compressed(GOLD / f"train_{tag}.npz", **gold)
    (GOLD / "train_pin_report.json").write_text(json.dumps(report, indent=1))
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
