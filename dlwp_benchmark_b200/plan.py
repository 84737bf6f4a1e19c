"""Plan cache: immutable twiddle tables owned by the C library, keyed by shape and device."""
from __future__ import annotations

import ctypes
import math
import threading
from typing import Dict, Tuple

import torch

from . import _lib


class Plan:
    """Handle to an ``sb200_plan_t``.  ``ky0``/``My``/``Mx`` describe the retained block:
    rows ky = (ky0 + j) mod H, j < My; cols kx < Mx."""

    def __init__(self, device: torch.device, H: int, W: int, ky0: int, My: int, Mx: int,
                 scale_fwd: float, scale_inv: float):
        lib = _lib.load()
        self.device = torch.device(device)
        self.H, self.W, self.ky0, self.My, self.Mx = H, W, ky0, My, Mx
        self.scale_fwd, self.scale_inv = scale_fwd, scale_inv
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.sb200_plan_create(ctypes.byref(handle), H, W, ky0, My, Mx, scale_fwd, scale_inv),
                       "sb200_plan_create")
        self.handle = handle

    def __del__(self):  # best effort; tables are tiny
        try:
            if getattr(self, "handle", None):
                _lib.load().sb200_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_plans: Dict[Tuple, Plan] = {}
_plock = threading.Lock()


def get_plan(device, H: int, W: int, ky0: int, My: int, Mx: int, scale_fwd: float, scale_inv: float) -> Plan:
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.SpectralB200Error("spectral_b200 plans need a CUDA device (no CPU path exists)")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, H, W, ky0, My, Mx, float(scale_fwd), float(scale_inv))
    with _plock:
        p = _plans.get(key)
        if p is None:
            p = Plan(torch.device("cuda", idx), H, W, ky0, My, Mx, scale_fwd, scale_inv)
            _plans[key] = p
    return p


def fno_mode_block(H: int, W: int, n_modes_halved) -> Tuple[int, int, int]:
    """(ky0, My, Mx) for neuralop's fftshift-era slicing (even H): the retained shifted rows are
    [st//2, H + (-st//2)) with st = H - min(H, n0); shifted row r <-> frequency r - H//2."""
    My = min(H, int(n_modes_halved[0]))
    Mx = min(W // 2 + 1, int(n_modes_halved[1]))
    st = H - My
    lo = st // 2
    return lo - H // 2, My, Mx


def fno_plan(device, H: int, W: int, n_modes_halved) -> Plan:
    if H % 2:
        raise _lib.SpectralB200Error(
            f"H={H}: odd grid heights are not supported by the CUDA path (the reference's "
            "double-fftshift quirk for odd sizes is not reproduced)")
    ky0, My, Mx = fno_mode_block(H, W, n_modes_halved)
    return get_plan(device, H, W, ky0, My, Mx, 1.0 / (H * W), 1.0)


def afno_plan(device, h: int, w: int, frac: float) -> Plan:
    total = h // 2 + 1
    kept = int(total * frac)
    r0, r1 = max(total - kept, 0), min(total + kept, h)
    kc = min(kept, w // 2 + 1)
    s = 1.0 / math.sqrt(h * w)
    return get_plan(device, h, w, r0, r1 - r0, kc, s, s)
