"""Post-processing (src/utils.py:55-64): CPU tests pin the restated algorithm against scipy (the reference's own
dependency) and, when the reference tree is present, against src/utils.py itself; GPU tests hold the CUDA kernel to
bit-exact peak positions and bit-exact float32 confidences."""
import sys
import types
from pathlib import Path

import numpy as np
import pytest

from oracle import post_oracle as PO

REF = Path("/root/reference")
PARAMS = dict(gauss_sigma=3.0, height=0.2, distance=15)      # src/ball_action/constants.py:39-43


@pytest.mark.parametrize("n,seed", [(400, 0), (1500, 1), (997, 2), (40, 3), (3, 4), (1, 5)])
def test_restated_algorithm_equals_scipy(n, seed):
    x = PO.synthetic_raw_predictions(n, 2, seed, events_every=60)
    for c in range(2):
        assert PO.post_processing_restated([15], x[:, c], **PARAMS) == PO.post_processing([15], x[:, c], **PARAMS)


def test_restated_algorithm_other_parameters():
    x = PO.synthetic_raw_predictions(600, 1, 7, events_every=40)[:, 0]
    for sigma, height, distance in [(1.0, 0.05, 1), (0.5, 0.5, 3), (6.0, 0.1, 40), (3.0, 0.0, 2)]:
        assert PO.post_processing_restated([0], x, sigma, height, distance) == PO.post_processing([0], x, sigma, height, distance)


@pytest.mark.skipif(not REF.exists(), reason="reference tree only exists in the authoring container")
def test_oracle_equals_reference_utils():
    sys.modules.setdefault("cv2", types.ModuleType("cv2"))       # src/utils.py imports cv2 for an unrelated helper
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_utils", REF / "src/utils.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    x = PO.synthetic_raw_predictions(3000, 2, 11)
    for c in range(2):
        assert mod.post_processing(list(range(15, 3015)), x[:, c], **PARAMS) == PO.post_processing(list(range(15, 3015)), x[:, c], **PARAMS)


GOLD = Path(__file__).parent / "golden" / "post_processing.npz"     # written by oracle/make_post_golden.py from src/utils.py
GOLD_CASES = [(5000, 1), (997, 2), (40, 3)]


@pytest.mark.parametrize("n,seed", GOLD_CASES)
def test_oracle_matches_reference_generated_golden(n, seed):
    g = np.load(GOLD)
    x = PO.synthetic_raw_predictions(n, 2, seed)
    for c in range(2):
        idx, conf = PO.post_processing(list(range(15, 15 + n)), x[:, c], **PARAMS)
        assert idx == g[f"idx_{n}_{seed}_{c}"].tolist()
        assert np.asarray(conf, dtype=np.float32).tolist() == g[f"conf_{n}_{seed}_{c}"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", GOLD_CASES)
def test_cuda_post_processing_matches_reference_generated_golden(n, seed):
    import torch
    from ball_action_spotting_b200 import postprocess as PP
    g = np.load(GOLD)
    x = PO.synthetic_raw_predictions(n, 2, seed)
    for c in range(2):
        idx, conf = PP.post_processing(list(range(15, 15 + n)), x[:, c], **PARAMS)
        assert idx == g[f"idx_{n}_{seed}_{c}"].tolist()
        assert np.asarray(conf, dtype=np.float32).tolist() == g[f"conf_{n}_{seed}_{c}"].tolist()


def test_results_spotting_document():
    from ball_action_spotting_b200 import postprocess as PP
    acts = {1: {"PASS": ([30, 1500], [0.9, 0.5]), "DRIVE": ([1499], [0.7])}, 2: {"PASS": ([25], [0.3]), "DRIVE": ([], [])}}
    assert PP.spotting_results(acts, "game/x", 25.0) == PO.spotting_results(acts, "game/x", 25.0)
    doc = PP.spotting_results(acts, "game/x", 25.0)
    assert doc["predictions"][0] == {"gameTime": "1 - 00:01", "label": "PASS", "position": "1200", "half": "1", "confidence": "0.9"}
    w, r = PP.gaussian_weights(3.0)
    assert r == 12 and abs(w.sum() - 1.0) < 1e-15 and w[12] == w.max()


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", [(67470, 0), (5000, 1), (997, 2), (33, 3), (2, 4), (1, 5)])
def test_cuda_post_processing_bit_exact(n, seed):
    import torch
    from ball_action_spotting_b200 import postprocess as PP
    x = PO.synthetic_raw_predictions(n, 2, seed)
    got = PP.find_actions(torch.from_numpy(x).cuda(), **PARAMS)
    for c in range(2):
        ref_idx, ref_conf = PO.post_processing([0], x[:, c], **PARAMS)
        assert got[c][0].cpu().tolist() == ref_idx
        assert got[c][1].cpu().numpy().tolist() == ref_conf            # bit-exact float32
    assert PP.post_processing(list(range(15, 15 + n)), x[:, 1], **PARAMS) == PO.post_processing(list(range(15, 15 + n)), x[:, 1], **PARAMS)


@pytest.mark.gpu
def test_cuda_post_processing_parameters_and_formats(tmp_path):
    import torch
    from ball_action_spotting_b200 import postprocess as PP
    x = PO.synthetic_raw_predictions(4000, 2, 9, events_every=50)
    for sigma, height, distance in [(1.0, 0.05, 1), (0.5, 0.5, 3), (6.0, 0.1, 40)]:
        got = PP.find_actions(torch.from_numpy(x).cuda(), sigma, height, distance)
        for c in range(2):
            ref_idx, ref_conf = PO.post_processing([0], x[:, c], sigma, height, distance)
            assert got[c][0].cpu().tolist() == ref_idx and got[c][1].cpu().numpy().tolist() == ref_conf
    frame_indexes = list(range(15, 4015))
    acts = PP.raw_predictions_to_actions(frame_indexes, x, {"PASS": 0, "DRIVE": 1}, PARAMS)
    ref = {cls: PO.post_processing(frame_indexes, x[:, ci], **PARAMS) for cls, ci in {"PASS": 0, "DRIVE": 1}.items()}
    assert acts == ref
    path = PP.prepare_game_spotting_results({1: acts}, "league/game", tmp_path, 25.0, PARAMS)
    import json
    assert json.loads(path.read_text()) == PO.spotting_results({1: ref}, "league/game", 25.0)
    PP.save_raw_predictions(tmp_path / "1_raw_predictions.npz", frame_indexes, torch.from_numpy(x))
    with np.load(tmp_path / "1_raw_predictions.npz") as z:
        assert z["frame_indexes"].tolist() == frame_indexes and np.array_equal(z["raw_predictions"], x)
