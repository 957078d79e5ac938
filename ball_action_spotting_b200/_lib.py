"""ctypes binding of libmds_b200.so (include/mds_b200.h).  There is no fallback: if the library is missing
or a call fails, a RuntimeError is raised (mirrors the reference's assert/raise convention)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "libmds_b200.so"


class MdsConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("num_classes", "num_frames", "stack_size", "num_3d_blocks", "num_3d_features",
                                       "num_3d_stack_proj", "expansion_3d_ratio", "se_reduce_3d_ratio", "device",
                                       "chunk_images")]


class MdsFrames(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int), ("img_stride", C.c_longlong), ("plane_stride", C.c_longlong),
                ("stored_h", C.c_int), ("pad_top", C.c_int), ("H", C.c_int), ("W", C.c_int), ("hflip", C.c_int)]


class MdsTrainConfig(C.Structure):
    _fields_ = ([(n, C.c_int) for n in ("num_classes", "num_frames", "stack_size", "num_3d_blocks", "num_3d_features",
                                        "num_3d_stack_proj", "expansion_3d_ratio", "se_reduce_3d_ratio", "device", "amp",
                                        "nesterov")] +
                [(n, C.c_float) for n in ("drop_rate", "drop_path_rate", "focal_alpha", "focal_gamma", "momentum",
                                          "init_scale")])


class MdsTrainStepArgs(C.Structure):
    _fields_ = [("enc_feats", C.c_void_p), ("targets", C.c_void_p), ("dp_masks", C.c_void_p), ("dropout_mask", C.c_void_p),
                ("seed", C.c_ulonglong), ("b", C.c_int), ("fh", C.c_int), ("fw", C.c_int), ("lr", C.c_float),
                ("apply_update", C.c_int), ("loss_out", C.c_void_p), ("logits_out", C.c_void_p)]


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_FP = C.POINTER(MdsFrames)

# name -> (restype, argtypes); must list every symbol declared in include/mds_b200.h
SIGNATURES = {
    "mds_last_error": (C.c_char_p, []),
    "mds_create": (_i, [C.POINTER(MdsConfig), C.POINTER(_vp)]),
    "mds_destroy": (_i, [_vp]),
    "mds_weights_add": (_i, [_vp, C.c_char_p, _vp, _sz]),
    "mds_weights_commit": (_i, [_vp]),
    "mds_workspace_bytes": (_sz, [_vp, _i, _i, _i, _i]),
    "mds_forward_2d": (_i, [_vp, _FP, _i, _vp, _vp, _sz, _vp]),
    "mds_forward_encoder": (_i, [_vp, _FP, _i, _vp, _vp, _sz, _vp]),
    "mds_forward_3d": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "mds_forward_head": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "mds_forward": (_i, [_vp, _FP, _i, _vp, _i, _vp, _sz, _vp]),
    "mds_nchw32_to_nhwc16": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mds_nhwc16_to_nchw32": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mds_gather_stacks": (_i, [_vp, _vp, C.c_longlong, _i, _i, _i, C.c_longlong, _vp]),
    "mds_axpby": (_i, [_vp, _vp, _f, _f, C.c_longlong, _vp]),
    "mds_k_stem": (_i, [_FP, _i, _vp, _vp, _vp, _vp]),
    "mds_k_conv3x3": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "mds_k_gemm1x1": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "mds_k_dwconv": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "mds_k_se_fc": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "mds_k_gemm_gated": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "mds_k_dwconv_se": (_i, [_vp] * 12 + [_i] * 9 + [_vp]),
    "mds_k_gemm_gate": (_i, [_vp] * 6 + [_i] * 5 + [_vp]),
    "mds_k_mbconv_tail": (_i, [_vp] * 15 + [_i] * 10 + [_vp]),
    "mds_k_gem": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _f, _vp]),
    "mds_k_linear": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mds_train_create": (_i, [C.POINTER(MdsTrainConfig), C.POINTER(_vp)]),
    "mds_train_destroy": (_i, [_vp]),
    "mds_train_num_tensors": (_i, [_vp, _i]),
    "mds_train_tensor_info": (_i, [_vp, _i, _i, C.POINTER(C.c_char_p), C.POINTER(C.c_longlong)]),
    "mds_train_set": (_i, [_vp, C.c_char_p, _vp, C.c_longlong]),
    "mds_train_get": (_i, [_vp, C.c_char_p, _i, _vp, C.c_longlong]),
    "mds_train_commit": (_i, [_vp, _vp]),
    "mds_train_ema_update": (_i, [_vp, C.c_double, _vp]),
    "mds_train_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "mds_train_step": (_i, [_vp, C.POINTER(MdsTrainStepArgs), _vp, _sz, _vp]),
    "mds_train_scaler_state": (_i, [_vp, _vp]),
    "mds_train_batches_tracked": (C.c_longlong, [_vp]),
    "mds_focal_loss": (_i, [_vp, _vp, _i, _f, _f, _vp, _vp, _vp]),
    "mds_post_processing_workspace_bytes": (_sz, [_i, _i]),
    "mds_post_processing": (_i, [_vp, _i, _i, _vp, _i, _f, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mds_set_pdl": (_i, [_i]),
    "mds_set_tail_mode": (_i, [_i]),
    "mds_set_streams": (_i, [_i]),
    "mds_set_conv_mode": (_i, [_i]),
    "mds_set_tail_dw_only": (_i, [_i]),
    "mds_launch_count": (C.c_longlong, [_i]),
    "mds_profile_begin": (_i, []),
    "mds_profile_end": (_i, [_vp, _vp, _vp, _i, _vp]),
}

_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """dlopen the in-tree library; raises if it has not been built (no CPU / torch fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA extension is mandatory, there is no fallback path)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mds_last_error()
        raise RuntimeError(f"libmds_b200 {what} failed ({rc}): {msg.decode() if msg else ''}")
