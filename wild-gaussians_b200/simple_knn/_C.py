"""``simple_knn._C`` of the reference (submodules/simple-knn/ext.cpp: one function, ``distCUDA2``) on this repo's kernels.

``distCUDA2(points)`` (spatial.cu:15-26): ``points`` is a float32 CUDA tensor [P,3]; returns a float32 tensor [P] with the
mean squared distance of every point to its three nearest other points.  Same values as the reference (exact search, the
reference's fp32 distance expression and summation order); there is no CPU fallback."""
from __future__ import annotations

import torch

from diff_gaussian_rasterization import _C as _gsr

__all__ = ["distCUDA2"]


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not isinstance(points, torch.Tensor) or not points.is_cuda:
        raise RuntimeError("distCUDA2: points must be a CUDA tensor (there is no CPU path, like the reference)")
    if points.dim() != 2 or points.size(1) != 3:
        raise RuntimeError("distCUDA2: points must have shape [P, 3]")
    pts = points.contiguous()
    if pts.dtype != torch.float32:          # the reference reinterprets the data as float (spatial.cu:23): require it
        raise RuntimeError("distCUDA2: points must be float32")
    P = int(pts.size(0))
    out = torch.zeros((P,), dtype=torch.float32, device=pts.device)      # torch::full({P}, 0.0) (spatial.cu:21)
    if P == 0:
        return out
    with torch.cuda.device(pts.device):
        nbytes = int(_gsr._lib.gsr_knn_scratch_bytes(P))
        scratch = torch.empty((nbytes + 256,), dtype=torch.uint8, device=pts.device)
        off = (-scratch.data_ptr()) % 256
        _gsr._check(_gsr._lib.gsr_knn_mean_dist2(P, pts.data_ptr(), out.data_ptr(), scratch.data_ptr() + off,
                                                 torch.cuda.current_stream(pts.device).cuda_stream), "gsr_knn_mean_dist2")
    return out
