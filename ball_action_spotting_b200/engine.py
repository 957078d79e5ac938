"""Thin Python wrapper over the C-ABI handle (include/mds_b200.h).  PyTorch is used only for device memory
and streams; every compute step is a call into libmds_b200.so."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import MdsConfig, MdsFrames, check


@dataclass(frozen=True)
class EngineConfig:
    num_classes: int = 2
    num_frames: int = 15
    stack_size: int = 3
    num_3d_blocks: int = 4
    num_3d_features: int = 192
    num_3d_stack_proj: int = 256
    expansion_3d_ratio: int = 3
    se_reduce_3d_ratio: int = 24
    chunk_images: int = 0

    @property
    def num_stacks(self) -> int:
        return self.num_frames // self.stack_size


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Engine:
    """One handle per (model, device).  Not thread-safe, like the reference predictor."""

    def __init__(self, cfg: EngineConfig, packed: Dict[str, torch.Tensor], device: torch.device):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ball_action_spotting_b200 runs on CUDA devices only (no CPU fallback)")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        c = MdsConfig(cfg.num_classes, cfg.num_frames, cfg.stack_size, cfg.num_3d_blocks, cfg.num_3d_features,
                      cfg.num_3d_stack_proj, cfg.expansion_3d_ratio, cfg.se_reduce_3d_ratio, idx, cfg.chunk_images)
        h = C.c_void_p()
        check(self.lib.mds_create(C.byref(c), C.byref(h)), "mds_create")
        self._h = h
        self._ws: Optional[torch.Tensor] = None
        self.version = 0          # bumped by every load_packed(): captured CUDA graphs of older versions must be rebuilt
        self.load_packed(packed)

    def load_packed(self, packed: Dict[str, torch.Tensor]) -> None:
        self.version += 1
        for name, t in packed.items():
            t = t.contiguous().cpu()
            check(self.lib.mds_weights_add(self._h, name.encode(), t.data_ptr(), t.numel() * t.element_size()),
                  f"mds_weights_add({name})")
        check(self.lib.mds_weights_commit(self._h), "mds_weights_commit")

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.mds_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- workspace -----------------------------------------------------------------------------------------
    def workspace(self, H: int, W: int, n_images: int, n_stacks: int) -> torch.Tensor:
        need = self.lib.mds_workspace_bytes(self._h, H, W, n_images, n_stacks)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    # ---- frames descriptor ---------------------------------------------------------------------------------
    @staticmethod
    def frames_desc(frames: torch.Tensor, H: int, W: int, img_stride: int, plane_stride: int, hflip: bool = False,
                    offset_elems: int = 0) -> MdsFrames:
        """frames: uint8 (raw, zero-padded to H in the kernel, frames.py:12-31) or float32 (already padded)."""
        if frames.dtype == torch.uint8:
            dtype = 0
        elif frames.dtype == torch.float32:
            dtype = 1
        else:
            raise RuntimeError(f"frames must be uint8 or float32, got {frames.dtype}")
        stored_h, stored_w = frames.shape[-2], frames.shape[-1]
        if stored_w != W:
            raise RuntimeError(f"frame width {stored_w} != {W}: width padding is not supported")
        if stored_h > H or (dtype == 1 and stored_h != H):
            raise RuntimeError(f"frame height {stored_h} incompatible with H={H}")
        pad_top = (H - stored_h) // 2
        return MdsFrames(frames.data_ptr() + offset_elems * frames.element_size(), dtype, img_stride, plane_stride,
                         stored_h, pad_top, H, W, 1 if hflip else 0)

    # ---- passes (fp16 NHWC in / out) -------------------------------------------------------------------------
    def forward_2d(self, desc: MdsFrames, n_images: int, feats_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        fh, fw = desc.H // 32, desc.W // 32
        if feats_out is None:
            feats_out = torch.empty((n_images, fh, fw, 192), dtype=torch.float16, device=self.device)
        ws = self.workspace(desc.H, desc.W, n_images, 0)
        check(self.lib.mds_forward_2d(self._h, C.byref(desc), n_images, feats_out.data_ptr(), ws.data_ptr(), ws.numel(),
                                      _stream(self.device)), "mds_forward_2d")
        return feats_out

    def forward_encoder(self, desc: MdsFrames, n_images: int) -> torch.Tensor:
        """conv2d_encoder(x)[-1] alone (multidim_stacker.py:215): fp16 (n_images, fh, fw, 192), input of conv2d_projection."""
        fh, fw = desc.H // 32, desc.W // 32
        feats = torch.empty((n_images, fh, fw, 192), dtype=torch.float16, device=self.device)
        ws = self.workspace(desc.H, desc.W, n_images, 0)
        check(self.lib.mds_forward_encoder(self._h, C.byref(desc), n_images, feats.data_ptr(), ws.data_ptr(), ws.numel(),
                                           _stream(self.device)), "mds_forward_encoder")
        return feats

    def forward_3d(self, feats: torch.Tensor) -> torch.Tensor:
        """feats fp16 (b, T, fh, fw, 192) contiguous -> fp16 (b, T, fh, fw, proj)."""
        b, T, fh, fw, c = feats.shape
        if T != self.cfg.num_stacks or c != self.cfg.num_3d_features:
            raise RuntimeError(f"forward_3d: expected (b, {self.cfg.num_stacks}, h, w, {self.cfg.num_3d_features}), got {tuple(feats.shape)}")
        out = torch.empty((b, T, fh, fw, self.cfg.num_3d_stack_proj), dtype=torch.float16, device=self.device)
        ws = self.workspace(fh * 32, fw * 32, 0, b)
        check(self.lib.mds_forward_3d(self._h, feats.data_ptr(), b, fh, fw, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                      _stream(self.device)), "mds_forward_3d")
        return out

    def forward_head(self, x: torch.Tensor, sigmoid: bool = False) -> torch.Tensor:
        """x fp16 (b, T, fh, fw, proj) -> float32 (b, num_classes)."""
        b, T, fh, fw, _ = x.shape
        out = torch.empty((b, self.cfg.num_classes), dtype=torch.float32, device=self.device)
        ws = self.workspace(fh * 32, fw * 32, 0, b)
        check(self.lib.mds_forward_head(self._h, x.data_ptr(), b, fh * fw, out.data_ptr(), 1 if sigmoid else 0,
                                        ws.data_ptr(), ws.numel(), _stream(self.device)), "mds_forward_head")
        return out

    def forward(self, desc: MdsFrames, b: int, sigmoid: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((b, self.cfg.num_classes), dtype=torch.float32, device=self.device)
        ws = self.workspace(desc.H, desc.W, b * self.cfg.num_stacks, b)
        check(self.lib.mds_forward(self._h, C.byref(desc), b, out.data_ptr(), 1 if sigmoid else 0, ws.data_ptr(),
                                   ws.numel(), _stream(self.device)), "mds_forward")
        return out

    # ---- boundary converters -------------------------------------------------------------------------------
    def nchw32_to_nhwc16(self, x: torch.Tensor) -> torch.Tensor:
        """float32 (n, C, *spatial) -> fp16 (n, *spatial, C)."""
        x = x.contiguous()
        n, c = x.shape[0], x.shape[1]
        sp = tuple(x.shape[2:])
        P = 1
        for s in sp:
            P *= s
        out = torch.empty((n, *sp, c), dtype=torch.float16, device=self.device)
        check(self.lib.mds_nchw32_to_nhwc16(x.data_ptr(), out.data_ptr(), n, c, P, _stream(self.device)), "nchw32_to_nhwc16")
        return out

    def nhwc16_to_nchw32(self, x: torch.Tensor) -> torch.Tensor:
        """fp16 (n, *spatial, C) -> float32 (n, C, *spatial)."""
        x = x.contiguous()
        n, c = x.shape[0], x.shape[-1]
        sp = tuple(x.shape[1:-1])
        P = 1
        for s in sp:
            P *= s
        out = torch.empty((n, c, *sp), dtype=torch.float32, device=self.device)
        check(self.lib.mds_nhwc16_to_nchw32(x.data_ptr(), out.data_ptr(), n, c, P, _stream(self.device)), "nhwc16_to_nchw32")
        return out

    def gather_stacks(self, feats: torch.Tensor, first_image: int, hop: int, n_pred: int, T: int) -> torch.Tensor:
        """feats fp16 (n_img, fh, fw, C) -> (n_pred, T, fh, fw, C): window p = images first_image + p + hop * t."""
        n_img = feats.shape[0]
        if first_image < 0 or first_image + (n_pred - 1) + hop * (T - 1) >= n_img:
            raise RuntimeError("gather_stacks: window reaches outside the cached features")
        plane = feats[0].numel()
        out = torch.empty((n_pred, T, *feats.shape[1:]), dtype=feats.dtype, device=self.device)
        for p0 in range(0, n_pred, 32768):
            n = min(32768, n_pred - p0)
            check(self.lib.mds_gather_stacks(feats.data_ptr(), out[p0:].data_ptr(), first_image + p0, hop, n, T, plane,
                                             _stream(self.device)), "mds_gather_stacks")
        return out

    def axpby_(self, y: torch.Tensor, x: torch.Tensor, a: float, b: float) -> torch.Tensor:
        """y <- a * y + b * x (float32, in place)."""
        check(self.lib.mds_axpby(y.data_ptr(), x.data_ptr(), float(a), float(b), y.numel(), _stream(self.device)), "mds_axpby")
        return y

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.mds_launch_count(1 if reset else 0))

    # ---- per-launch CUDA-event profile (bench.py) ------------------------------------------------------------
    def profile_begin(self) -> None:
        check(self.lib.mds_profile_begin(), "mds_profile_begin")

    def profile_end(self, capacity: int = 1 << 16):
        """-> list of (kind, tag, ms) for every kernel launched since profile_begin()."""
        kinds = (C.c_int * capacity)()
        tags = (C.c_int * capacity)()
        ms = (C.c_float * capacity)()
        count = C.c_int(0)
        check(self.lib.mds_profile_end(kinds, tags, ms, capacity, C.byref(count)), "mds_profile_end")
        n = min(count.value, capacity)
        return [(kinds[i], tags[i], ms[i]) for i in range(n)]
