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


def _load_checkpoint(model_path) -> dict:
    """torch.load restricted to tensors and plain containers (the EmaCheckpoint dict holds nothing else); a checkpoint
    that needs arbitrary unpickling is only loaded when MDS_UNSAFE_LOAD=1."""
    import os
    import pickle
    try:
        return torch.load(str(model_path), map_location="cpu", weights_only=True)
    except pickle.UnpicklingError:
        if os.environ.get("MDS_UNSAFE_LOAD") == "1":
            return torch.load(str(model_path), map_location="cpu", weights_only=False)
        raise


def load_model(model_path, device: str = "cuda:0", bias_correction: bool = True) -> LoadedModel:
    """Reads the ``EmaCheckpoint`` dict (src/ema.py:71-76): {model_name, params, nn_state_dict, ...}.
    bias_correction: packer.py's data-free correction of the fp16 weight-rounding bias (off = the exact folded parameters)."""
    state = _load_checkpoint(model_path)
    params = state["params"]
    name, kwargs = params["nn_module"]
    assert name == "multidim_stacker"                                     # predictors.py:26
    kwargs = dict(kwargs)
    kwargs["pretrained"] = False
    module = MultiDimStacker(**kwargs, bias_correction=bias_correction)
    module.load_state_dict(state["nn_state_dict"])
    dev = torch.device(device[0] if isinstance(device, (list, tuple)) else device)
    module.to(dev).eval()
    return LoadedModel(params, module, dev)


class _FrameGraphs:
    """CUDA graphs of the two fixed-shape pieces of one ``predict`` call: (1) encoder of the newest 3-frame triple,
    (2) 3D blocks + head over the window of cached features.  One call launches ~90 kernels of a few microseconds each and
    is bound by the host (0.97 ms of launch + Python work per frame vs 1.09 ms total), so the launch sequence is recorded
    once and replayed; inputs and outputs live at fixed addresses."""

    def __init__(self, predictor: "MultiDimStackerPredictor", h: int, w: int):
        dev = predictor.device
        eng = predictor.model.nn_module.engine(dev)
        W, H = predictor.image_size
        n_tta = 2 if predictor.tta else 1
        T = predictor.frame_stack_size // predictor.model_stack_size
        self.frames = torch.zeros((predictor.model_stack_size, h, w), dtype=torch.uint8, device=dev)
        self.stack = torch.zeros((n_tta, T, H // 32, W // 32, 192), dtype=torch.float16, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up outside capture: workspaces, kernel attributes
            for _ in range(2):
                predictor._encode(eng, self.frames)
                predictor._head(eng, self.stack)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.g2d = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g2d):
            self.feat = predictor._encode(eng, self.frames)              # (n_tta, fh, fw, 192) fp16
        self.g3d = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g3d):
            self.pred = predictor._head(eng, self.stack)                 # (num_classes,) float32
        self._ws = eng._ws          # the recorded launches point into this workspace: keep it alive if the engine grows a new one
        self.version = eng.version  # weights (and gem_p, a kernel argument) are baked into the recording

    def encode(self, frames) -> torch.Tensor:
        for i, f in enumerate(frames):
            self.frames[i].copy_(f)
        self.g2d.replay()
        return self.feat.clone()

    def head(self, feats) -> torch.Tensor:
        for j, f in enumerate(feats):
            self.stack[:, j].copy_(f)
        self.g3d.replay()
        return self.pred.clone()


class MultiDimStackerPredictor:
    def __init__(self, model_path: Path, device: str = "cuda:0", tta: bool = False, cuda_graph: bool = True,
                 bias_correction: bool = True):
        self.model = load_model(model_path, device=device, bias_correction=bias_correction)
        self.cuda_graph = cuda_graph
        self._graphs: Optional[_FrameGraphs] = None
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

    def _encode(self, eng, frames: torch.Tensor) -> torch.Tensor:
        """frames (3, h, W) uint8 -> cached features (n_tta, fh, fw, 192) fp16 (predictors.py:60-66; the TTA flip is the
        stem kernel reading mirrored columns, :63)."""
        h, w = frames.shape[-2:]
        W, H = self.image_size
        outs = []
        for flip in ((False, True) if self.tta else (False,)):
            desc = eng.frames_desc(frames, H, W, 3 * h * w, h * w, hflip=flip)
            outs.append(eng.forward_2d(desc, 1))                           # (1, fh, fw, 192) fp16
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    def _head(self, eng, feats: torch.Tensor) -> torch.Tensor:
        """feats (n_tta, T, fh, fw, 192) -> probabilities (num_classes,) (predictors.py:69-72)."""
        prediction = eng.forward_head(eng.forward_3d(feats), sigmoid=True)  # prediction_transform fused
        if prediction.shape[0] == 2:                                       # TTA: mean of the two branches (:72)
            return eng.axpby_(prediction[0], prediction[1], 0.5, 0.5)
        return prediction[0]

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
            eng = self.model.nn_module.engine(self.device)
            if self.cuda_graph and (self._graphs is None or self._graphs.frames.shape[-2:] != frame.shape[-2:]
                                    or self._graphs.version != eng.version):        # re-packed weights invalidate the recording
                self._graphs = _FrameGraphs(self, frame.shape[-2], frame.shape[-1])
            stacks_indexes = list(batched(predict_indexes, self.model_stack_size))
            for stack_indexes in stacks_indexes:
                if stack_indexes not in self._stack_indexes2features:
                    triple = [self._frame_index2frame[i] for i in stack_indexes]
                    self._stack_indexes2features[stack_indexes] = (
                        self._graphs.encode(triple) if self.cuda_graph else self._encode(eng, torch.stack(triple, dim=0)))
            cached = [self._stack_indexes2features[s] for s in stacks_indexes]     # T x (n_tta, fh, fw, 192)
            if self.cuda_graph:
                return self._graphs.head(cached), predict_index
            return self._head(eng, torch.stack(cached, dim=1).contiguous()), predict_index
        return None, predict_index
