"""TEST INFRASTRUCTURE -- binding of ``oracle/_ref/libdgr_ref.so`` (the UNMODIFIED reference CUDA
rasterizer, built by ``oracle/Makefile`` from ``/root/reference``; see ``oracle/ref_shim.cu``).

Exposes the reference's ``_C`` surface (``rasterize_gaussians`` / ``rasterize_gaussians_backward`` /
``mark_visible``, ``rasterize_points.cu:35-225``) with the same positional arguments and returned
tuples, reproducing what ``rasterize_points.cu`` does around the core: output allocation, zero-filled
gradient tensors, byte buffers grown through the three resize callbacks.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import this module.  The product
(``wild-gaussians_b200/``) never does.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import CFUNCTYPE, POINTER, c_char_p, c_float, c_int, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdgr_ref.so")

_ALLOC = CFUNCTYPE(c_void_p, c_void_p, c_int, c_size_t)
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(LIB_PATH)
        lib.ref_last_error.restype = c_char_p
        lib.ref_forward.restype = c_int
        lib.ref_forward.argtypes = [_ALLOC, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int] + \
            [c_void_p] * 5 + [c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                              c_void_p, c_int, c_void_p, c_void_p, c_int]
        lib.ref_backward.restype = c_int
        lib.ref_backward.argtypes = [c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p] + [c_void_p] * 9 + [c_int]
        lib.ref_mark_visible.restype = c_int
        lib.ref_mark_visible.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        _lib = lib
    return _lib


def _ptr(t):
    return None if (t is None or t.numel() == 0) else t.data_ptr()


def _c(t, dev):
    if t.numel() == 0:
        return t
    return t.to(device=dev, dtype=torch.float32).contiguous()


def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        viewmatrix, projmatrix, tan_fovx, tan_fovy, kernel_size, subpixel_offset, image_height,
                        image_width, sh, degree, campos, prefiltered, debug):
    lib = _load()
    dev = means3D.device
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    out_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    bufs = [torch.empty((0,), dtype=torch.uint8, device=dev) for _ in range(3)]  # geom, binning, img
    if P == 0:
        return 0, out_color, radii, bufs[0], bufs[1], bufs[2]
    M = int(sh.size(1)) if sh.numel() != 0 else 0
    keep = [_c(t, dev) for t in (background, means3D, colors, opacity, scales, rotations, cov3D_precomp, viewmatrix,
                                 projmatrix, subpixel_offset, sh, campos)]
    (background, means3D, colors, opacity, scales, rotations, cov3D_precomp, viewmatrix, projmatrix,
     subpixel_offset, sh, campos) = keep

    def alloc(_ctx, which, nbytes):
        bufs[which] = torch.empty((int(nbytes),), dtype=torch.uint8, device=dev)
        return bufs[which].data_ptr()

    cb = _ALLOC(alloc)
    with torch.cuda.device(dev):
        R = lib.ref_forward(cb, None, P, int(degree), M, _ptr(background), W, H, _ptr(means3D), _ptr(sh),
                            _ptr(colors), _ptr(opacity), _ptr(scales), float(scale_modifier), _ptr(rotations),
                            _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), float(tan_fovx),
                            float(tan_fovy), float(kernel_size), _ptr(subpixel_offset), int(bool(prefiltered)),
                            out_color.data_ptr(), radii.data_ptr(), int(bool(debug)))
    if R < 0:
        raise RuntimeError("reference forward failed: " + lib.ref_last_error().decode())
    return R, out_color, radii, bufs[0], bufs[1], bufs[2]


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, kernel_size,
                                 subpixel_offset, dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer,
                                 imageBuffer, debug):
    lib = _load()
    dev = means3D.device
    P = int(means3D.size(0))
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    M = int(sh.size(1)) if sh.numel() != 0 else 0
    z = dict(dtype=torch.float32, device=dev)
    dL_dmeans3D = torch.zeros((P, 3), **z)
    dL_dmeans2D = torch.zeros((P, 3), **z)
    dL_dcolors = torch.zeros((P, 3), **z)
    dL_dconic = torch.zeros((P, 2, 2), **z)
    dL_dopacity = torch.zeros((P, 1), **z)
    dL_dcov3D = torch.zeros((P, 6), **z)
    dL_dsh = torch.zeros((P, M, 3), **z)
    dL_dscales = torch.zeros((P, 3), **z)
    dL_drotations = torch.zeros((P, 4), **z)
    if P != 0:
        keep = [_c(t, dev) for t in (background, means3D, colors, scales, rotations, cov3D_precomp, viewmatrix,
                                     projmatrix, subpixel_offset, dL_dout_color, sh, campos)]
        (background, means3D, colors, scales, rotations, cov3D_precomp, viewmatrix, projmatrix, subpixel_offset,
         dL_dout_color, sh, campos) = keep
        with torch.cuda.device(dev):
            rc = lib.ref_backward(P, int(degree), M, int(R), _ptr(background), W, H, _ptr(means3D), _ptr(sh),
                                  _ptr(colors), _ptr(scales), float(scale_modifier), _ptr(rotations),
                                  _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                                  float(tan_fovx), float(tan_fovy), float(kernel_size), _ptr(subpixel_offset),
                                  radii.data_ptr(), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                                  _ptr(dL_dout_color), dL_dmeans2D.data_ptr(), dL_dconic.data_ptr(),
                                  dL_dopacity.data_ptr(), dL_dcolors.data_ptr(), dL_dmeans3D.data_ptr(),
                                  dL_dcov3D.data_ptr(), _ptr(dL_dsh), dL_dscales.data_ptr(),
                                  dL_drotations.data_ptr(), int(bool(debug)))
        if rc != 0:
            raise RuntimeError("reference backward failed: " + lib.ref_last_error().decode())
    return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)


def mark_visible(means3D, viewmatrix, projmatrix):
    lib = _load()
    P = int(means3D.size(0))
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P:
        lib.ref_mark_visible(P, means3D.contiguous().data_ptr(), viewmatrix.contiguous().data_ptr(),
                             projmatrix.contiguous().data_ptr(), present.data_ptr())
    return present


# ---- views into the reference's opaque buffers (layouts: rasterizer_impl.cu:155-194) -------------

def _align(x, a=128):
    return (x + a - 1) // a * a


def debug_views(geomBuffer, binningBuffer, imgBuffer, P, W, H, R):
    """Integer artefacts of the reference, for bit-exact comparison.

    Geometry chunk order: depths f32[P], clamped bool[3P], internal_radii i32[P], means2D float2[P],
    cov3D f32[6P], conic_opacity float4[P], rgb f32[3P], tiles_touched u32[P], scan space, point_offsets.
    Binning: point_list u32[R], point_list_unsorted u32[R], keys u64[R], keys_unsorted u64[R].
    Image: accum_alpha f32[N], n_contrib u32[N], ranges uint2[N].
    """
    out = {}
    N = W * H

    def take(buf, cur, nbytes):
        start = _align(buf.data_ptr() + cur) - buf.data_ptr()
        return buf[start:start + nbytes], start + nbytes

    cur = 0
    v, cur = take(geomBuffer, cur, 4 * P); out["depths"] = v.view(torch.float32)
    v, cur = take(geomBuffer, cur, 3 * P)
    v, cur = take(geomBuffer, cur, 4 * P)
    v, cur = take(geomBuffer, cur, 8 * P); out["means2D"] = v.view(torch.float32).view(P, 2)
    v, cur = take(geomBuffer, cur, 24 * P); out["cov3D"] = v.view(torch.float32).view(P, 6)
    v, cur = take(geomBuffer, cur, 16 * P); out["conic_opacity"] = v.view(torch.float32).view(P, 4)
    v, cur = take(geomBuffer, cur, 12 * P); out["rgb"] = v.view(torch.float32).view(P, 3)
    v, cur = take(geomBuffer, cur, 4 * P); out["tiles_touched"] = v.view(torch.int32)

    cur = 0
    v, cur = take(imgBuffer, cur, 4 * N); out["final_T"] = v.view(torch.float32).view(H, W)
    v, cur = take(imgBuffer, cur, 4 * N); out["n_contrib"] = v.view(torch.int32).view(H, W)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    v, cur = take(imgBuffer, cur, 8 * N); out["ranges"] = v.view(torch.int32).view(N, 2)[:T]

    if R > 0:
        cur = 0
        v, cur = take(binningBuffer, cur, 4 * R); out["point_list"] = v.view(torch.int32)
    else:
        out["point_list"] = torch.empty((0,), dtype=torch.int32, device=imgBuffer.device)
    return out
