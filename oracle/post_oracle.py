"""CPU oracle for the post-processing step.  TEST INFRASTRUCTURE ONLY (imported by tests/ only).

``post_processing`` below is ``src/utils.py:55-64`` of the reference verbatim in behaviour: it calls the reference's own
third-party dependencies, ``scipy.ndimage.gaussian_filter`` and ``scipy.signal.find_peaks`` (requirements.txt pins
scipy==1.10.1; this image has scipy 1.18 — the two functions' documented algorithms are unchanged: 'reflect' boundary,
truncate 4.0, plateau midpoints, greedy minimal-distance selection by peak height).  ``spotting_results`` restates
``prepare_game_spotting_results`` (``src/ball_action/annotations.py:88-111``) without touching the disk.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import gaussian_filter
from scipy.signal import find_peaks


def post_processing(frame_indexes, predictions: np.ndarray, gauss_sigma: float, height: float, distance: int):
    predictions = gaussian_filter(predictions, gauss_sigma)
    peaks, _ = find_peaks(predictions, height=height, distance=distance)
    confidences = predictions[peaks].tolist()
    action_frame_indexes = (peaks + frame_indexes[0]).tolist()
    return action_frame_indexes, confidences


def spotting_results(half2class_actions: dict, game: str, video_fps: float = 25.0) -> dict:
    results = {"UrlLocal": game, "predictions": []}
    for half in half2class_actions.keys():
        for cls, (frame_indexes, confidences) in half2class_actions[half].items():
            for frame_index, confidence in zip(frame_indexes, confidences):
                position = round(frame_index / video_fps * 1000)
                seconds = int(frame_index / video_fps)
                results["predictions"].append({"gameTime": f"{half} - {seconds // 60:02}:{seconds % 60:02}", "label": cls,
                                               "position": str(position), "half": str(half), "confidence": str(confidence)})
    results["predictions"] = sorted(results["predictions"], key=lambda pred: (int(pred["half"]), int(pred["position"])))
    return results


def synthetic_raw_predictions(n: int, k: int = 2, seed: int = 0, events_every: int = 180) -> np.ndarray:
    """Probabilities that look like the model's output: mostly ~0 with bumps of random width / height, plus exact
    plateaus and ties (saturated 1.0 runs, repeated bumps) to exercise the plateau and tie rules."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, k)).astype(np.float32) * 0.05
    t = np.arange(n, dtype=np.float32)
    for c in range(k):
        pos = rng.integers(0, n, size=max(1, n // events_every))
        for p in pos:
            width, amp = rng.uniform(1.5, 8.0), rng.uniform(0.1, 1.0)
            x[:, c] += (amp * np.exp(-0.5 * ((t - p) / width) ** 2)).astype(np.float32)
        for p in rng.integers(0, max(1, n - 40), size=max(1, n // (4 * events_every))):
            x[p:p + rng.integers(2, 30), c] = 1.0          # saturated run -> plateau after clipping
    return np.clip(x, 0.0, 1.0).astype(np.float32)


def post_processing_restated(frame_indexes, predictions: np.ndarray, gauss_sigma: float, height: float, distance: int):
    """The same result computed the way the CUDA kernel does (csrc/postproc.cuh), in plain numpy / Python loops: used by
    the CPU tests to check, against scipy, the published algorithms the kernel restates -- NI_Correlate1D's symmetric
    summation in double with 'reflect' indexing, _local_maxima_1d's plateau midpoints, and _select_by_peak_distance's
    greedy selection evaluated in parallel rounds."""
    x = np.asarray(predictions, dtype=np.float32)
    n = x.shape[0]
    radius = int(4.0 * float(gauss_sigma) + 0.5)
    k = np.arange(-radius, radius + 1)
    w = np.exp(-0.5 / (gauss_sigma * gauss_sigma) * k ** 2)
    w = w / w.sum()

    def refl(i):
        if n == 1:
            return 0
        i %= 2 * n
        return i if i < n else 2 * n - 1 - i
    sm = np.empty(n, dtype=np.float32)
    xd = x.astype(np.float64)
    for i in range(n):
        acc = xd[i] * w[radius]
        for j in range(-radius, 0):
            acc += (xd[refl(i + j)] + xd[refl(i - j)]) * w[radius + j]
        sm[i] = np.float32(acc)
    state = np.zeros(n, dtype=np.int8)
    for i in range(1, n - 1):
        if not sm[i - 1] < sm[i]:
            continue
        ahead = i + 1
        while ahead < n - 1 and sm[ahead] == sm[i]:
            ahead += 1
        if sm[ahead] < sm[i]:
            mid = (i + ahead - 1) // 2
            if sm[mid] >= np.float32(height):
                state[mid] = 1
    if distance > 1:
        while (state == 1).any():
            top = []
            for i in np.nonzero(state == 1)[0]:
                ok = True
                for d in range(1, distance):
                    l, r = i - d, i + d
                    if l >= 0 and state[l] == 1 and sm[l] > sm[i]:
                        ok = False
                    if r < n and state[r] == 1 and sm[r] >= sm[i]:
                        ok = False
                if ok:
                    top.append(i)
            for i in top:
                state[i] = 2
            for i in np.nonzero(state == 1)[0]:
                lo, hi = max(0, i - distance + 1), min(n, i + distance)
                if (state[lo:hi] == 2).any():
                    state[i] = 3
    else:
        state[state == 1] = 2
    peaks = np.nonzero(state == 2)[0]
    return (peaks + frame_indexes[0]).tolist(), sm[peaks].tolist()
