"""ctypes binding of libgifgan.so (include/gifgan.h).

This is the only place the product touches the CUDA kernels.  There is NO CPU
fallback: if the shared library is missing, or a tensor is not on a CUDA device,
the call raises.  PyTorch only supplies device memory and the current stream.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GG_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libgifgan.so")   # GG_LIB: A/B builds of the same library

GG_F32, GG_BF16 = 0, 1
ACT = {None: 0, "none": 0, "relu": 1, "lrelu": 2, "tanh": 3, "sigmoid": 4, "tanh01": 5}
CONV_TENSOR_CORE = 2


class ConvDesc(C.Structure):
    """struct gg_conv_desc (include/gifgan.h)."""
    _fields_ = [(n, C.c_int32) for n in
                ("N", "D", "H", "W", "C", "Do", "Ho", "Wo", "K", "kd", "kh", "kw", "sd", "sh", "sw", "pd", "ph", "pw",
                 "large_dtype", "small_dtype", "act")] + [("act_param", C.c_float), ("flags", C.c_int32)]


_lib = None

_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
_dp = C.POINTER(ConvDesc)

# name -> (restype, argtypes); mirrors include/gifgan.h one to one
SIGNATURES = {
    "gg_version": (C.c_int, []),
    "gg_last_error": (C.c_char_p, []),
    "gg_device_arch": (C.c_int, []),
    "gg_launch_count": (C.c_uint64, []),
    "gg_debug_set_repeat": (None, [C.c_int]),
    "gg_debug_set_prof": (None, [C.c_void_p]),
    "gg_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "gg_ipc_import": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "gg_dp_signal_bytes": (C.c_size_t, []),
    "gg_dp_allreduce": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _i32, _i32, _i64, _i64, _i32, _vp]),
    "gg_workspace_bytes": (C.c_size_t, []),
    "gg_set_workspace": (C.c_int, [C.c_void_p, C.c_size_t]),
    "gg_conv_down": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _vp]),
    "gg_conv_up": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _vp]),
    "gg_conv_wgrad": (C.c_int, [_dp, _vp, _vp, _vp, _vp]),
    "gg_conv_wgrad_bias": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _vp]),
    "gg_conv_down_stats": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "gg_conv_up_stats": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "gg_conv_dgrad_bnbwd": (C.c_int, [_dp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _i32, _vp, C.POINTER(C.c_int32), _vp]),
    "gg_bn_fwd_train_stats": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _i32, _f32, _vp, _vp]),
    "gg_conv2d_fwd": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _vp]),
    "gg_conv2d_dgrad": (C.c_int, [_dp, _vp, _vp, _vp, _vp]),
    "gg_conv2d_wgrad": (C.c_int, [_dp, _vp, _vp, _vp, _vp]),
    "gg_deconv2d_fwd": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _vp]),
    "gg_deconv2d_dgrad": (C.c_int, [_dp, _vp, _vp, _vp, _vp]),
    "gg_deconv2d_wgrad": (C.c_int, [_dp, _vp, _vp, _vp, _vp]),
    "gg_conv3d_fwd": (C.c_int, [_dp, _vp, _vp, _vp, _vp, _vp]),
    "gg_conv3d_dgrad": (C.c_int, [_dp, _vp, _vp, _vp, _vp]),
    "gg_conv3d_wgrad": (C.c_int, [_dp, _vp, _vp, _vp, _vp]),
    "gg_linear_fwd": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "gg_linear_fwd_stats": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "gg_linear_dgrad": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "gg_linear_wgrad": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp]),
    "gg_linear_bn_ok": (C.c_int, [_i32, _i32, _i32, _i32, _i32]),
    "gg_linear_bn_fwd": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _i32, _f32, _vp]),
    "gg_linear_bn_bwd": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _vp]),
    "gg_bn_workspace_bytes": (_sz, [_i32, _i32]),
    "gg_bn_fwd_train": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _i32, _f32, _vp, _sz, _vp]),
    "gg_bn_fwd_infer": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _f32, _i32, _f32, _vp]),
    "gg_bn_bwd": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _i32, _vp, _sz, _vp]),
    "gg_bn_infer_stats": (C.c_int, [_vp, _vp, _f32, _i32, _vp, _vp, _vp]),
    "gg_act_bwd": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _i64, _i32, _f32, _vp]),
    "gg_act_fwd": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _i32, _f32, _vp]),
    "gg_act_bwd_bias": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i32, _i64, _i32, _i32, _f32, _vp, _vp]),
    "gg_bias_grad": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _vp]),
    "gg_cast": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _vp]),
    "gg_axpby": (C.c_int, [_vp, _f32, _vp, _f32, _i64, _vp]),
    "gg_gather_scalars": (C.c_int, [C.POINTER(C.c_void_p), _i32, _vp, _vp]),
    "gg_get_std": (C.c_int, [_vp, _i32, _i64, _i64, _vp, _vp, _sz, _vp]),
    "gg_sigmoid_ce": (C.c_int, [_vp, _i64, _f32, _f32, _vp, _i32, _vp, _vp]),
    "gg_mse": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i64, _f32, _vp, _i32, _vp, _vp]),
    "gg_frames_to_input": (C.c_int, [_vp, _i32, _i32, _i32, _i64, _i64, _vp, _i32, _i32, _i32, _vp]),
    "gg_loss_head_ok": (C.c_int, [_i32, _i32, _i32]),
    "gg_loss_head_fwd": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "gg_loss_head_bwd": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _i32, _i32, _vp, _vp, _vp]),
    "gg_distance_loss_workspace_bytes": (_sz, []),
    "gg_distance_loss": (C.c_int, [_vp, _i32, _vp, _i64, _f32, _f32, _vp, _i32, _vp, _vp, _sz, _vp]),
    "gg_adam": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _vp]),
    "gg_adam_graph": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _f32, _f32, _vp]),
    "gg_adam_tick": (C.c_int, [_vp, _f32, _f32, _f32, _vp]),
    "gg_adam_apply": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _f32, _f32, _f32, _f32, _vp]),
    "gg_lstm_step_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp]),
    "gg_lstm_step_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp]),
}


def lib():
    """Load libgifgan.so (once).  Fails loudly: there is no other implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(gif-gan_b200 has no CPU or PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.gg_version() != 100:
            raise RuntimeError("libgifgan.so version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().gg_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libgifgan {what} failed (status {rc}): {msg}")


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return GG_F32
    if t.dtype == torch.bfloat16:
        return GG_BF16
    raise TypeError(f"unsupported dtype {t.dtype}: the B200 kernels take float32 or bfloat16")


def torch_dtype(code: int):
    return torch.float32 if code == GG_F32 else torch.bfloat16


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("gif-gan_b200 kernels run on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("non-contiguous tensor passed to libgifgan")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def gather_scalars(tensors, dst):
    """dst[i] = tensors[i].reshape(-1)[0] for up to 8 device tensors (None entries are skipped): one launch."""
    n = len(tensors)
    arr = (C.c_void_p * n)(*[None if t is None else t.data_ptr() for t in tensors])
    check(lib().gg_gather_scalars(arr, n, ptr(dst), stream()), "gg_gather_scalars")


def launch_count() -> int:
    return int(lib().gg_launch_count())
