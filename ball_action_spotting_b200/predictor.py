"""Drop-in for ``src/predictors.py::MultiDimStackerPredictor`` (boundary B2).

Same constructor, attributes (``device``, ``indexes_generator``, ``model``, ``tta``) and streaming contract:
``predict(frame, index) -> (Optional[Tensor(num_classes,)], predict_index)``; ``None`` until the whole window
is buffered (predictors.py:57,74-75).  The per-triple 2D-feature cache of the reference (:59-67) is kept, but it
holds fp16 NHWC engine tensors, frames stay uint8 on the device (padding + /255 are fused into the stem
kernel), and the flip of the TTA branch is done by the stem kernel instead of a kornia copy.
"""
from __future__ import annotations

from itertools import islice
from pathlib import Path
from typing import Iterable, Optional

import torch
from torch import nn

from .indexes import StackIndexesGenerator
from .model import MultiDimStacker


def batched(iterable: Iterable, size: int):
    iterator = iter(iterable)
    while batch := tuple(islice(iterator, size)):
        yield batch


class LoadedModel:
    """What ``argus.load_model`` returns, reduced to the attributes the predictor and scripts read
    (SURVEY.md Appendix B): nn_module, params, device, prediction_transform, eval()."""

    def __init__(self, params: dict, nn_module: MultiDimStacker, device: torch.device):
        self.params, self.nn_module, self.device = params, nn_module, device
        self.prediction_transform = nn.Sigmoid()

    def eval(self):
        self.nn_module.eval()
        return self

    def get_nn_module(self):
        return self.nn_module


def load_model(model_path, device: str = "cuda:0") -> LoadedModel:
    """Reads the ``EmaCheckpoint`` dict (src/ema.py:71-76): {model_name, params, nn_state_dict, ...}."""
    state = torch.load(str(model_path), map_location="cpu", weights_only=False)
    params = state["params"]
    name, kwargs = params["nn_module"]
    assert name == "multidim_stacker"                                     # predictors.py:26
    kwargs = dict(kwargs)
    kwargs["pretrained"] = False
    module = MultiDimStacker(**kwargs)
    module.load_state_dict(state["nn_state_dict"])
    dev = torch.device(device[0] if isinstance(device, (list, tuple)) else device)
    module.to(dev).eval()
    return LoadedModel(params, module, dev)


class MultiDimStackerPredictor:
    def __init__(self, model_path: Path, device: str = "cuda:0", tta: bool = False):
        self.model = load_model(model_path, device=device)
        self.model.eval()
        self.device = self.model.device
        self.tta = tta
        assert self.model.params["nn_module"][0] == "multidim_stacker"
        fp_name, fp_params = self.model.params["frames_processor"]
        assert fp_name == "pad_normalize" and fp_params.get("pad_mode", "constant") == "constant" \
            and fp_params.get("fill_value", 0) == 0, "only the constant-0 pad_normalize processor is fused"
        self.image_size = tuple(fp_params["size"])                         # (W, H)
        self.frame_stack_size = self.model.params["frame_stack_size"]
        self.frame_stack_step = self.model.params["frame_stack_step"]
        self.indexes_generator = StackIndexesGenerator(self.frame_stack_size, self.frame_stack_step)
        self.model_stack_size = self.model.params["nn_module"][1]["stack_size"]

        self._frame_index2frame: dict = dict()
        self._stack_indexes2features: dict = dict()
        self._predict_offset: int = self.indexes_generator.make_stack_indexes(0)[-1]

    def reset_buffers(self):
        self._frame_index2frame = dict()
        self._stack_indexes2features = dict()

    def _clear_old(self, minimum_index: int):
        for index in [i for i in self._frame_index2frame if i < minimum_index]:
            del self._frame_index2frame[index]
        for stack_indexes in [s for s in self._stack_indexes2features if any(i < minimum_index for i in s)]:
            del self._stack_indexes2features[stack_indexes]

    def _features_2d(self, stack_indexes) -> torch.Tensor:
        eng = self.model.nn_module.engine(self.device)
        frames = torch.stack([self._frame_index2frame[i] for i in stack_indexes], dim=0)   # (3, h, W) uint8
        h, w = frames.shape[-2:]
        W, H = self.image_size
        outs = []
        for flip in ((False, True) if self.tta else (False,)):
            desc = eng.frames_desc(frames, H, W, 3 * h * w, h * w, hflip=flip)
            outs.append(eng.forward_2d(desc, 1))                           # (1, fh, fw, 192) fp16
        return torch.cat(outs, dim=0)                                      # (B_tta, fh, fw, 192)

    @torch.no_grad()
    def predict(self, frame: torch.Tensor, index: int):
        frame = frame.to(device=self.device)
        if frame.dtype != torch.uint8:
            raise RuntimeError("predict() takes the raw uint8 frame (frames.py normalisation is fused on the GPU)")
        self._frame_index2frame[index] = frame.contiguous()
        predict_index = index - self._predict_offset
        predict_indexes = self.indexes_generator.make_stack_indexes(predict_index)
        self._clear_old(predict_indexes[0])
        if set(predict_indexes) <= set(self._frame_index2frame.keys()):
            stacks_indexes = list(batched(predict_indexes, self.model_stack_size))
            for stack_indexes in stacks_indexes:
                if stack_indexes not in self._stack_indexes2features:
                    self._stack_indexes2features[stack_indexes] = self._features_2d(stack_indexes)
            feats = torch.stack([self._stack_indexes2features[s] for s in stacks_indexes], dim=1)  # (B, T, fh, fw, 192)
            eng = self.model.nn_module.engine(self.device)
            x = eng.forward_3d(feats.contiguous())
            prediction = eng.forward_head(x, sigmoid=True)                 # prediction_transform fused
            if prediction.shape[0] == 2:                                   # TTA: mean of the two branches (predictors.py:72)
                prediction = eng.axpby_(prediction[0], prediction[1], 0.5, 0.5)
            else:
                prediction = prediction[0]
            return prediction, predict_index
        return None, predict_index
