"""Raw predictions -> action spots and the on-disk formats downstream tools read.

Mirrors ``post_processing`` (``src/utils.py:55-64``), ``raw_predictions_to_actions`` / ``prepare_game_spotting_results``
(``src/ball_action/annotations.py:76-115``) and the ``raw_predictions.npz`` layout of ``scripts/ball_action/predict.py:79-83``.
The arithmetic (gaussian_filter + find_peaks for every class) is one launch of ``mds_post_processing``; this module only
prepares the gaussian weights the way scipy does and formats the results.  CUDA only, no CPU path.
"""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check


def gaussian_weights(sigma: float, truncate: float = 4.0) -> Tuple[np.ndarray, int]:
    """scipy.ndimage ``_gaussian_kernel1d(sigma, 0, radius)`` with ``radius = int(truncate * sigma + 0.5)``."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return np.ascontiguousarray(phi / phi.sum(), dtype=np.float64), radius


def find_actions(raw_predictions: torch.Tensor, gauss_sigma: float, height: float, distance: int
                 ) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """raw_predictions: CUDA float32 (N, num_classes).  Per class: (peak positions int32, confidences float32), on device."""
    if not raw_predictions.is_cuda or raw_predictions.dtype != torch.float32 or raw_predictions.ndim != 2:
        raise RuntimeError("find_actions: raw_predictions must be a CUDA float32 (N, num_classes) tensor")
    lib = _lib.load()
    raw = raw_predictions.contiguous()
    n, k = raw.shape
    dev = raw.device
    if n == 0:
        return [(torch.empty(0, dtype=torch.int32, device=dev), torch.empty(0, dtype=torch.float32, device=dev)) for _ in range(k)]
    w, radius = gaussian_weights(gauss_sigma)
    idx = torch.empty((k, n), dtype=torch.int32, device=dev)
    conf = torch.empty((k, n), dtype=torch.float32, device=dev)
    count = torch.empty(k, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.mds_post_processing_workspace_bytes(n, k), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.mds_post_processing(raw.data_ptr(), n, k, w.ctypes.data_as(C.c_void_p), radius, float(height), int(distance),
                                      idx.data_ptr(), conf.data_ptr(), count.data_ptr(), ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream(dev).cuda_stream), "mds_post_processing")
    counts = count.cpu().tolist()
    return [(idx[c, :counts[c]], conf[c, :counts[c]]) for c in range(k)]


def post_processing(frame_indexes: Sequence[int], predictions, gauss_sigma: float, height: float, distance: int,
                    device: str = "cuda:0") -> Tuple[List[int], List[float]]:
    """Same signature and return value as ``src/utils.py:55-64`` (1-D ``predictions`` of one class)."""
    pred = torch.as_tensor(np.asarray(predictions, dtype=np.float32) if not torch.is_tensor(predictions) else predictions)
    pred = pred.to(device=device, dtype=torch.float32).reshape(-1, 1)
    (peaks, conf), = find_actions(pred, gauss_sigma, height, distance)
    return (peaks.cpu().numpy().astype(np.int64) + int(frame_indexes[0])).tolist(), conf.cpu().numpy().tolist()


def raw_predictions_to_actions(frame_indexes: Sequence[int], raw_predictions, class2target: Dict[str, int],
                               postprocess_params: Dict[str, float], device: str = "cuda:0") -> Dict[str, tuple]:
    """``src/ball_action/annotations.py:76-85``; all classes share one kernel launch."""
    raw = torch.as_tensor(raw_predictions).to(device=device, dtype=torch.float32)
    per_class = find_actions(raw, postprocess_params["gauss_sigma"], postprocess_params["height"], postprocess_params["distance"])
    out = {}
    for cls, ci in class2target.items():
        peaks, conf = per_class[ci]
        out[cls] = ((peaks.cpu().numpy().astype(np.int64) + int(frame_indexes[0])).tolist(), conf.cpu().numpy().tolist())
    return out


def save_raw_predictions(path, frame_indexes: Sequence[int], raw_predictions) -> None:
    """``{half}_raw_predictions.npz`` (scripts/ball_action/predict.py:79-83)."""
    raw = raw_predictions.detach().cpu().numpy() if torch.is_tensor(raw_predictions) else np.asarray(raw_predictions)
    np.savez(str(path), frame_indexes=np.asarray(list(frame_indexes)), raw_predictions=raw)


def spotting_results(half2class_actions: Dict[int, Dict[str, tuple]], game: str, video_fps: float) -> dict:
    """The ``results_spotting.json`` document of ``prepare_game_spotting_results`` (annotations.py:88-111)."""
    predictions = []
    for half, class_actions in half2class_actions.items():
        for cls, (frame_indexes, confidences) in class_actions.items():
            for frame_index, confidence in zip(frame_indexes, confidences):
                seconds = int(frame_index / video_fps)
                predictions.append({"gameTime": f"{half} - {seconds // 60:02}:{seconds % 60:02}", "label": cls,
                                    "position": str(round(frame_index / video_fps * 1000)), "half": str(half),
                                    "confidence": str(confidence)})
    predictions.sort(key=lambda pred: (int(pred["half"]), int(pred["position"])))
    return {"UrlLocal": game, "predictions": predictions}


def prepare_game_spotting_results(half2class_actions, game: str, prediction_dir, video_fps: float,
                                  postprocess_params: Dict[str, float]) -> Path:
    game_dir = Path(prediction_dir) / game
    game_dir.mkdir(parents=True, exist_ok=True)
    path = game_dir / "results_spotting.json"
    with open(path, "w") as f:
        json.dump(spotting_results(half2class_actions, game, video_fps), f, indent=4)
    with open(game_dir / "postprocess_params.json", "w") as f:
        json.dump(postprocess_params, f, indent=4)
    return path
