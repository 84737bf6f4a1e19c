"""ctypes binding of the C-ABI kernel library ``libspectral_b200.so`` (include/spectral_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
caller gets an exception (the product path must never silently run on the CPU / on torch.fft).
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspectral_b200.so")

_lib = None
_lock = threading.Lock()

_vp, _i, _i64, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double

# name -> (restype, argtypes); mirrors include/spectral_b200.h one to one
SIGNATURES = {
    "sb200_version": (_i, []),
    "sb200_last_error": (ctypes.c_char_p, []),
    "sb200_kernel_launches": (_i64, []),
    "sb200_device_arch": (_i, []),
    "sb200_plan_create": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _d, _d]),
    "sb200_plan_destroy": (_i, [_vp]),
    "sb200_rowdft_fwd": (_i, [_vp, _i, _vp, _vp, _i64, _vp, _i]),
    "sb200_coldft_fwd": (_i, [_vp, _i, _vp, _vp, _i64, _vp]),
    "sb200_coldft_inv": (_i, [_vp, _i, _vp, _vp, _i64, _vp]),
    "sb200_analysis_scratch": (_i64, [_vp, _i64]),
    "sb200_analysis": (_i, [_vp, _i, _vp, _vp, _i64, _vp, _vp, _i]),
    "sb200_modes_gemm": (_i, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i, _i, _i, _i, _i, _vp]),
    "sb200_mlp_head_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _vp, _i]),
    "sb200_mlp_head_bwd_workspace": (_i64, []),
    "sb200_mlp_head_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _vp, _i]),
    "sb200_lift_tail_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _vp, _i]),
    "sb200_lift_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _vp, _i]),
    "sb200_lift_wgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _vp, _i]),
    "sb200_cgemm_workspace": (_i64, [_vp, _i]),
    "sb200_cgemm_grouped": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "sb200_cgemm": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "sb200_rowidft_pointwise": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i]),
    "sb200_pointwise_wgrad_workspace": (_i64, [_i, _i, _i, _i64, _i]),
    "sb200_pointwise_wgrad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _vp, _vp, _i]),
    "sb200_pointwise_small_n": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i, _vp]),
    "sb200_wgrad_small_workspace": (_i64, [_i, _i, _i, _i64]),
    "sb200_wgrad_small": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i, _vp, _vp]),
    "sb200_cl_rowdft_fwd": (_i, [_vp, _i, _vp, _vp, _i64, _i, _vp]),
    "sb200_cl_coldft_fwd": (_i, [_vp, _i, _vp, _vp, _i, _i, _vp]),
    "sb200_cl_coldft_inv": (_i, [_vp, _i, _vp, _vp, _i, _i, _vp]),
    "sb200_cl_rowidft_res": (_i, [_vp, _i, _vp, _vp, _vp, _i64, _i, _vp]),
    "sb200_cl_rowidft_res2": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "sb200_afno_blocklinear_fwd": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, ctypes.c_float, _vp]),
    "sb200_afno_blocklinear_dgrad": (_i, [_vp, _vp, _i, _vp, _vp, _i64, _i, _i, _i, _vp]),
    "sb200_afno_blocklinear_wgrad_workspace": (_i64, [_i64, _i, _i, _i]),
    "sb200_afno_blocklinear_wgrad": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i64, _i, _i, _i, _vp, _vp]),
    "sb200_gelu_fwd": (_i, [_vp, _vp, _i64, _vp]),
    "sb200_gelu_bwd": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "sb200_gemm_workspace": (_i64, [_i, _i, _i, _i, _i, _i]),
    "sb200_gemm": (_i, [_vp, _i64, _i, _vp, _i64, _i, _vp, _i64, _i, _i, _i, _vp, _i, _vp, _i64, _vp, _i64, _i, _vp, _i64,
                        _i, _i, _i, _vp, _vp, _i]),
    "sb200_gemm_batched_workspace": (_i64, [_vp, _i]),
    "sb200_gemm_batched": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "sb200_afno_embed": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "sb200_afno_unembed": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "sb200_mask_mul": (_i, [_vp, _vp, _vp, _i64, _i, _vp]),
    "sb200_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, ctypes.c_float, _vp]),
    "sb200_layernorm_bwd_workspace": (_i64, [_i64, _i]),
    "sb200_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _vp]),
    "sb200_colsum_workspace": (_i64, [_i64, _i]),
    "sb200_colsum": (_i, [_vp, _i64, _vp, _i64, _i, _vp, _vp]),
    "sb200_batch_sum": (_i, [_vp, _vp, _i, _i64, _vp]),
    "sb200_channel_sum_workspace": (_i64, [_i, _i]),
    "sb200_channel_sum": (_i, [_vp, _vp, _i, _i, _i64, _vp, _vp]),
}


class SpectralB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise SpectralB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C dlwp_benchmark_b200/csrc`. There is no CPU / torch.fft fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


# Precision mode of the tensor-core stages (0 = CUDA-core fp32, 1 = single-pass TF32, 3 = 3xTF32 parity mode): host-side
# state, handed to the library as an argument of every call (the library itself is stateless).
_tc_mode = 3


def set_tc_mode(mode: int) -> int:
    """Set the precision mode used by every later call (forward and backward); returns the previous mode."""
    global _tc_mode
    if mode not in (0, 1, 3):
        raise SpectralB200Error(f"tc mode must be 0, 1 or 3, got {mode!r}")
    prev, _tc_mode = _tc_mode, int(mode)
    return prev


def tc_mode() -> int:
    return _tc_mode


def on_tensor_device(fn):
    """Decorator for ``autograd.Function.forward/backward``: run with the CUDA device of the first CUDA tensor
    argument current, so that ``torch.cuda.current_stream()``, the plan cache, workspaces allocated with
    ``torch.empty`` and the library's per-device lookups all refer to the tensor's device even when the caller's
    current device is another one (the backward runs on an autograd engine thread)."""
    import functools

    @functools.wraps(fn)
    def wrapper(ctx, *args):
        import torch
        dev = next((a.device for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if dev is None:
            return fn(ctx, *args)
        with torch.cuda.device(dev):
            return fn(ctx, *args)
    return wrapper


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().sb200_last_error().decode("utf-8", "replace")
        raise SpectralB200Error(f"{what or 'spectral_b200'} failed (rc={rc}): {msg}")
