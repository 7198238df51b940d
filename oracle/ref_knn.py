"""TEST INFRASTRUCTURE -- ctypes loader of the UNMODIFIED reference simple-knn compiled by ``make -C oracle knn_ref``
(oracle/_ref/libsimpleknn_ref.so; sources stay under /root/reference).  ``distCUDA2`` mirrors spatial.cu:15-26."""
from __future__ import annotations

import ctypes
import os

import torch

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libsimpleknn_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH) and torch.cuda.is_available()


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
        _lib.ref_knn_mean_dist2.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib.ref_knn_mean_dist2.restype = ctypes.c_int
    pts = points.contiguous().float()
    out = torch.zeros((pts.shape[0],), dtype=torch.float32, device=pts.device)
    torch.cuda.synchronize(pts.device)          # the reference runs on the legacy default stream
    with torch.cuda.device(pts.device):
        rc = _lib.ref_knn_mean_dist2(int(pts.shape[0]), pts.data_ptr(), out.data_ptr())
    if rc != 0:
        raise RuntimeError(f"reference SimpleKNN::knn failed: cuda error {rc}")
    return out
