"""Training step with a frozen 2D encoder (BASELINE.json configs[4], ``configs/ball_action/ball_finetune_long_004.py``).

Host-side mirror of ``BallActionModel.train_step`` (``src/argus_models.py:41-74``) for configs that set
``freeze_conv2d_encoder`` (``:104-110``): the EfficientNetV2 encoder runs through the inference engine (its
parameters get no gradient), everything after it — ``conv2d_projection``, the ``InvertedResidual3d`` blocks,
``conv3d_projection``, GeM, dropout, classifier, focal loss, backward, GradScaler, SGD-Nesterov — is one call into
``libmds_b200.so`` (``mds_train_step``).  PyTorch only supplies device memory and the stream.

Deliberate difference from the reference: the reference calls ``self.train()`` on the whole module, so its *frozen*
encoder still normalises with batch statistics, updates its BatchNorm running statistics and applies DropPath
(SURVEY.md §3.4).  Here the frozen encoder is evaluated in eval mode (running statistics, no DropPath), i.e. it
really is frozen; the trainable part follows the reference exactly (train-mode BatchNorm, DropPath, Dropout).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import MdsTrainConfig, MdsTrainStepArgs, check
from .model import MultiDimStacker


class FrozenEncoderTrainer:
    """One trainer per (model, device); owns fp32 master weights, gradients, momentum and BN statistics on the GPU.

    ``loss`` / ``optimizer`` follow the argus params of the reference config
    (``configs/ball_action/ball_finetune_long_004.py:46-55``): focal loss (alpha, gamma, mean reduction) and
    SGD (lr, momentum, nesterov).
    """

    def __init__(self, model: MultiDimStacker, lr: float, momentum: float = 0.9, nesterov: bool = True,
                 focal_alpha: float = 0.4, focal_gamma: float = 1.2, amp: bool = True,
                 drop_rate: Optional[float] = None, drop_path_rate: Optional[float] = None, init_scale: float = 65536.0,
                 ema_decay: Optional[float] = None, batch_transform=None):
        """batch_transform: optional callable (frames, target) -> (frames, target) applied on the device before the encoder,
        where the reference applies its GPU augmentations and mixup (argus_models.py:49-53)."""
        self.lib = _lib.load()
        self.batch_transform = batch_transform
        self.model = model
        dev = model.classifier.weight.device
        if dev.type != "cuda":
            raise RuntimeError("FrozenEncoderTrainer: the model must live on a CUDA device (no CPU path)")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        cfg = model._cfg
        self.lr = float(lr)
        self.ema_decay = ema_decay        # ModelEma decay (configs: ema_decay=0.999); None = no averaging
        self.focal_alpha, self.focal_gamma = float(focal_alpha), float(focal_gamma)
        self._val_module: Optional[MultiDimStacker] = None
        self._val_version = -1
        self._ema_tracked = None
        self.cfg = MdsTrainConfig(
            cfg.num_classes, cfg.num_frames, cfg.stack_size, cfg.num_3d_blocks, cfg.num_3d_features, cfg.num_3d_stack_proj,
            cfg.expansion_3d_ratio, cfg.se_reduce_3d_ratio, self.device.index, 1 if amp else 0, 1 if nesterov else 0,
            float(model.drop_rate if drop_rate is None else drop_rate),
            float(model.drop_path_rate if drop_path_rate is None else drop_path_rate), float(focal_alpha),
            float(focal_gamma), float(momentum), float(init_scale))
        h = C.c_void_p()
        check(self.lib.mds_train_create(C.byref(self.cfg), C.byref(h)), "mds_train_create")
        self._h = h
        self._ws: Optional[torch.Tensor] = None
        self._step = 0
        self.param_names = self._names(0)
        self.buffer_names = self._names(1)
        self._tracked_base = int(model.state_dict()[next(k for k in model.state_dict() if k.startswith('conv3d_projection') and k.endswith('num_batches_tracked'))])
        self.load_from_module()

    def _names(self, kind: int) -> Dict[str, int]:
        out = {}
        for i in range(self.lib.mds_train_num_tensors(self._h, kind)):
            name, numel = C.c_char_p(), C.c_longlong()
            check(self.lib.mds_train_tensor_info(self._h, kind, i, C.byref(name), C.byref(numel)), "mds_train_tensor_info")
            out[name.value.decode()] = numel.value
        return out

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.mds_train_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters <-> nn.Module ---------------------------------------------------------------------------------
    def load_from_module(self) -> None:
        sd = self.model.state_dict()
        for name, numel in {**self.param_names, **self.buffer_names}.items():
            t = sd[name].detach().float().cpu().contiguous()
            if t.numel() != numel:
                raise RuntimeError(f"{name}: expected {numel} elements, state_dict has {t.numel()}")
            check(self.lib.mds_train_set(self._h, name.encode(), t.data_ptr(), numel), f"mds_train_set({name})")
        check(self.lib.mds_train_commit(self._h, torch.cuda.current_stream(self.device).cuda_stream), "mds_train_commit")

    def get(self, name: str, what: str = "value") -> torch.Tensor:
        """what: 'value' | 'grad' (of the last step, unscaled) | 'momentum' | 'ema'.  Flat float32 CPU tensor."""
        numel = self.param_names.get(name, self.buffer_names.get(name))
        if numel is None:
            raise KeyError(name)
        out = torch.empty(numel, dtype=torch.float32)
        check(self.lib.mds_train_get(self._h, name.encode(), {"value": 0, "grad": 1, "momentum": 2, "ema": 3}[what], out.data_ptr(), numel),
              f"mds_train_get({name})")
        return out

    def sync_to_module(self) -> None:
        """Write the trained parameters and BatchNorm statistics back into the nn.Module (checkpointing / inference)."""
        sd = self.model.state_dict()
        tracked = int(self.lib.mds_train_batches_tracked(self._h))
        with torch.no_grad():
            for name in list(self.param_names) + list(self.buffer_names):
                sd[name].copy_(self.get(name).view(sd[name].shape))
            for name in self.buffer_names:
                if name.endswith("running_mean"):
                    k = name[: -len("running_mean")] + "num_batches_tracked"
                    sd[k].add_(tracked - getattr(self, "_tracked_synced", 0))
        self._tracked_synced = tracked
        self.model.repack()

    def ema_update(self) -> None:
        """``self.model_ema.update(self.nn_module)`` (argus_models.py:68-69, src/ema.py:49-57)."""
        if self.ema_decay is None:
            raise RuntimeError("ema_update: the trainer was created without ema_decay")
        check(self.lib.mds_train_ema_update(self._h, float(self.ema_decay), torch.cuda.current_stream(self.device).cuda_stream),
              "mds_train_ema_update")
        # num_batches_tracked is an integer buffer: ModelEma averages it in float and copy_ truncates (src/ema.py:47)
        tracked = int(self.lib.mds_train_batches_tracked(self._h)) + self._tracked_base
        if self._ema_tracked is None:
            self._ema_tracked = self._tracked_base
        self._ema_tracked = int((torch.tensor(float(self.ema_decay)) * self._ema_tracked + torch.tensor(1.0 - float(self.ema_decay)) * tracked).item())

    def ema_state_dict(self) -> Dict[str, torch.Tensor]:
        """state_dict of ``model_ema.ema`` (what EmaCheckpoint stores as ``nn_state_dict``, src/ema.py:61-76): averaged trainable
        parameters and BatchNorm statistics; the frozen encoder is constant, so its entries equal the module's."""
        sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        for name in list(self.param_names) + list(self.buffer_names):
            sd[name] = self.get(name, "ema").view(sd[name].shape).to(sd[name].device)
        if self._ema_tracked is not None:
            for name in self.buffer_names:
                if name.endswith("running_mean"):
                    k = name[: -len("running_mean")] + "num_batches_tracked"
                    sd[k] = torch.tensor(self._ema_tracked, dtype=torch.long, device=sd[k].device)
        return sd

    #: how this trainer deviates from ``BallActionModel.train_step``; ``checkpoint_params`` records it next to the weights
    DEVIATIONS = {"b200_frozen_encoder_mode": "eval (running BatchNorm statistics, no DropPath); the reference keeps the frozen "
                                              "encoder in train mode (argus_models.py:42,104-110)",
                  "b200_iter_size": 1}

    def checkpoint_params(self, params: dict) -> dict:
        """``params`` for ``EmaCheckpoint`` (src/ema.py:71-76) with the deviations of this trainer recorded, so that a checkpoint
        trained here is distinguishable from a reference-trained one (its conv2d_encoder BatchNorm buffers stay at their
        pre-trained values instead of drifting with the batch statistics)."""
        out = dict(params)
        out.update(self.DEVIATIONS)
        return out

    def scaler_state(self) -> Tuple[float, float, float, float]:
        out = (C.c_float * 4)()
        check(self.lib.mds_train_scaler_state(self._h, out), "mds_train_scaler_state")
        return tuple(out)

    # ---- the step ---------------------------------------------------------------------------------------------------
    def _workspace(self, b: int, fh: int, fw: int) -> torch.Tensor:
        need = self.lib.mds_train_workspace_bytes(self._h, b, fh, fw)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def encoder_features(self, frames: torch.Tensor) -> torch.Tensor:
        """(b, num_frames, H, W) uint8 / float32 frames -> fp16 (b, T, H/32, W/32, 192) frozen-encoder output."""
        was_training = self.model.training
        self.model.eval()
        try:
            eng = self.model.engine(frames.device)
        finally:
            self.model.train(was_training)
        b, t, h, w = frames.shape
        frames = frames.contiguous()
        T = t // self.model.stack_size
        desc = eng.frames_desc(frames, self.model._padded_h(frames, h), w, self.model.stack_size * h * w, h * w)
        return eng.forward_encoder(desc, b * T).view(b, T, desc.H // 32, desc.W // 32, 192)

    def step_on_features(self, enc_feats: torch.Tensor, targets: torch.Tensor, dp_masks: Optional[torch.Tensor] = None,
                         dropout_mask: Optional[torch.Tensor] = None, apply_update: bool = True,
                         lr: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """enc_feats fp16 (b, T, fh, fw, 192); targets f32 (b, classes) -> (loss (1,), logits (b, classes)), on device."""
        if enc_feats.dtype != torch.float16 or not enc_feats.is_cuda:
            raise RuntimeError("step_on_features: enc_feats must be a CUDA fp16 tensor")
        b, T, fh, fw, c = enc_feats.shape
        if T != self.cfg.num_frames // self.cfg.stack_size or c != 192:
            raise RuntimeError(f"step_on_features: bad feature shape {tuple(enc_feats.shape)}")
        enc_feats = enc_feats.contiguous()
        targets = targets.to(self.device, torch.float32).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        logits = torch.empty((b, self.cfg.num_classes), dtype=torch.float32, device=self.device)
        keep = []

        def opt(t):
            if t is None:
                return None
            t = t.to(self.device, torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()
        self._step += 1
        args = MdsTrainStepArgs(enc_feats.data_ptr(), targets.data_ptr(), opt(dp_masks), opt(dropout_mask), self._step,
                                b, fh, fw, float(self.lr if lr is None else lr), 1 if apply_update else 0,
                                loss.data_ptr(), logits.data_ptr())
        ws = self._workspace(b, fh, fw)
        check(self.lib.mds_train_step(self._h, C.byref(args), ws.data_ptr(), ws.numel(),
                                      torch.cuda.current_stream(self.device).cuda_stream), "mds_train_step")
        return loss, logits

    def train_step(self, batch, state=None) -> dict:
        """``BallActionModel.train_step(batch, state)`` (argus_models.py:41-74): batch = (frames, target)."""
        frames, target = batch
        frames = frames.to(self.device, non_blocking=True)
        target = target.to(self.device, non_blocking=True)
        if self.batch_transform is not None:                 # GPU augmentations / mixup of the reference (argus_models.py:49-53)
            with torch.no_grad():
                frames, target = self.batch_transform(frames, target)
        loss, logits = self.step_on_features(self.encoder_features(frames), target)
        if self.ema_decay is not None:
            self.ema_update()
        return {"prediction": torch.sigmoid(logits), "target": target, "loss": loss.item()}     # prediction_transform (:26,69)

    def val_step(self, batch, state=None) -> dict:
        """``BallActionModel.val_step(batch, state)`` (argus_models.py:76-91): eval-mode forward of the EMA model when
        averaging is on (else of the current weights), focal loss, sigmoid.  The evaluation copy of the module is refreshed
        only when a training step has happened since the last call."""
        frames, target = batch
        frames = frames.to(self.device, non_blocking=True)
        target = target.to(self.device, torch.float32, non_blocking=True).contiguous()
        if self._val_module is None:          # a second module (its own engine handle) holding the weights to evaluate
            c = self.model._cfg
            self._val_module = MultiDimStacker("tf_efficientnetv2_b0.in1k", c.num_classes, num_frames=c.num_frames,
                                               stack_size=c.stack_size, num_3d_blocks=c.num_3d_blocks,
                                               num_3d_features=c.num_3d_features, num_3d_stack_proj=c.num_3d_stack_proj,
                                               expansion_3d_ratio=c.expansion_3d_ratio, se_reduce_3d_ratio=c.se_reduce_3d_ratio,
                                               drop_rate=self.model.drop_rate, chunk_images=c.chunk_images,
                                               bias_correction=self.model._bias_correction).to(self.device).eval()
        if self._val_version != self._step:
            if self.ema_decay is not None:
                self._val_module.load_state_dict(self.ema_state_dict())
            else:
                self.sync_to_module()
                self._val_module.load_state_dict(self.model.state_dict())
            self._val_version = self._step
        logits = self._val_module(frames).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        probs = torch.empty_like(logits)
        check(self.lib.mds_focal_loss(logits.data_ptr(), target.data_ptr(), logits.numel(), self.focal_alpha, self.focal_gamma,
                                      loss.data_ptr(), probs.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream),
              "mds_focal_loss")
        return {"prediction": probs, "target": target, "loss": loss.item()}
