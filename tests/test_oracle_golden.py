"""CPU: pin the oracle against the committed golden fixtures (tests/golden, written by oracle/make_golden.py after
checking bit-exact equality with the unmodified reference module), and — when /root/reference is present (authoring
container) — against the reference module itself."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import mds_oracle as O

GOLD = Path(__file__).parent / "golden"
REF = Path("/root/reference")


def _run(cfg, hw, tap=None):
    sd = O.make_state_dict(cfg, seed=1234)
    x = torch.rand((1, cfg.num_frames, *hw), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        f2 = O.forward_2d(sd, x, cfg, tap)
        f3 = O.forward_3d(sd, f2, cfg, tap)
        lg = O.forward_head(sd, f3, tap)
    return f2, f3, lg


@pytest.mark.parametrize("tag,cfg,hw", [("t5_small", O.ModelConfig(), (96, 160)),
                                        ("t11_small", O.ModelConfig(num_frames=33), (96, 160))])
def test_oracle_matches_golden(tag, cfg, hw):
    g = np.load(GOLD / f"oracle_{tag}.npz")
    taps = {}
    f2, f3, lg = _run(cfg, hw, lambda n, t: taps.__setitem__(n, t))
    # golden values were produced on another CPU: allow reduction-order noise only (1e-4 relative)
    np.testing.assert_allclose(lg.numpy(), g["logits"], rtol=1e-4, atol=1e-4 * np.abs(g["logits"]).max())
    np.testing.assert_allclose(f2.flatten()[:64].numpy(), g["forward_2d_head"], rtol=0, atol=1e-4 * float(g["forward_2d_absmax"]))
    np.testing.assert_allclose(f3.flatten()[:64].numpy(), g["forward_3d_head"], rtol=0, atol=1e-4 * float(g["forward_3d_absmax"]))
    for k in g.files:
        if k.startswith("tap_absmax."):
            name = k.split(".", 1)[1]
            assert abs(float(taps[name].abs().max()) - float(g[k])) <= 1e-3 * float(g[k]) + 1e-6, name


def test_oracle_full_size_logits_golden():
    g = np.load(GOLD / "oracle_t5_full.npz")
    _, _, lg = _run(O.ModelConfig(), (736, 1280))
    np.testing.assert_allclose(lg.numpy(), g["logits"], rtol=1e-4, atol=1e-4 * np.abs(g["logits"]).max())


def test_oracle_u8_golden():
    g = np.load(GOLD / "oracle_u8_small.npz")
    cfg = O.ModelConfig()
    sd = O.make_state_dict(cfg, seed=1234)
    u8 = torch.randint(0, 256, (1, 15, 80, 160), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        lg = O.forward(sd, O.pad_normalize(u8, (160, 96)), cfg)
    np.testing.assert_allclose(lg.numpy(), g["logits"], rtol=1e-4, atol=1e-4 * np.abs(g["logits"]).max())


def test_pin_report_records_bit_exact_match_with_reference():
    rep = json.loads((GOLD / "oracle_pin_report.json").read_text())
    for k, v in rep.items():
        if k.endswith("max_abs_diff_vs_reference"):
            assert v == 0.0, k
    assert rep["t5_full.encoder_params"] == 5610384
    assert rep["torchvision.full_params_1000cls"] == 7139704      # timm's published 7.14 M for tf_efficientnetv2_b0
    assert rep["torchvision.encoder_rel_diff"] < 1e-5


@pytest.mark.skipif(not REF.exists(), reason="reference tree only exists in the authoring container")
def test_oracle_vs_unmodified_reference_module():
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "oracle"))
    import make_golden as MG
    cfg = O.ModelConfig()
    sd = O.make_state_dict(cfg, seed=77)
    ref = MG.build_reference(cfg, sd)
    x = torch.rand((2, 15, 64, 96), generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        a = ref(x)
        b = O.forward(sd, x, cfg)
    assert torch.equal(a, b)


def test_indexes_and_window_semantics():
    assert O.make_stack_indexes(20, 15, 2) == list(range(6, 35, 2))
    assert O.make_stack_indexes(0, 15, 2)[-1] == 14             # _predict_offset, predictors.py:36
    assert O.clip_index(0, 1000, 15, 2, 1) == 15 and O.clip_index(1000, 1000, 15, 2, 1) == 984
    assert O.stack_offsets(33, 2) == (32, 32)
