"""Seeded synthetic Gaussian clouds and cameras.

One generator for the parity tests, ``bench.py`` and the golden-vector script, so that every backend sees
identical bits: everything is drawn on the CPU with ``torch.Generator().manual_seed(seed)``.  The camera is
built exactly like ``wildgaussians/method.py:1502-1519`` builds it (OpenCV pinhole, ``znear=0.01``,
``zfar=100``, matrices transposed to the rasterizer's row-vector convention).
"""
from __future__ import annotations

import math

import torch

C0 = 0.28209479177387814


def projection_from_opencv(w, h, fx, fy, cx, cy, znear, zfar):
    """``getProjectionMatrixFromOpenCV`` (method.py:605-616)."""
    P = torch.zeros((4, 4))
    P[0, 0] = 2.0 * fx / w
    P[1, 1] = 2.0 * fy / h
    P[0, 2] = (2.0 * cx - w) / w
    P[1, 2] = (2.0 * cy - h) / h
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(W, H, world_view=None):
    fx = fy = 0.9 * W
    cx, cy = W / 2.0, H / 2.0
    wv = torch.eye(4) if world_view is None else world_view.clone().float()
    world_view_transform = wv.transpose(0, 1).contiguous()
    projection = projection_from_opencv(W, H, fx, fy, cx, cy, 0.01, 100.0).transpose(0, 1)
    full_proj = world_view_transform.unsqueeze(0).bmm(projection.unsqueeze(0)).squeeze(0).contiguous()
    campos = world_view_transform.inverse()[3, :3].contiguous()
    tanfovx = math.tan(2 * math.atan(W / (2 * fx)) * 0.5)
    tanfovy = math.tan(2 * math.atan(H / (2 * fy)) * 0.5)
    return dict(viewmatrix=world_view_transform, projmatrix=full_proj, campos=campos, tanfovx=tanfovx,
                tanfovy=tanfovy)


def make_scene(P, W, H, sh_degree=None, seed=0, scale_range=(0.002, 0.03), normalize_rot=True, bg=(0.0, 0.0, 0.0),
               subpixel_jitter=0.0, max_sh_degree=3, cov3D_precomp=False, kernel_size=0.1):
    """Returns a dict of CPU tensors / python scalars describing one rasterizer call.

    ``sh_degree=None``: colours given as ``colors_precomp`` (P,3), the path ``method.py`` uses.
    ``sh_degree=d``: ``shs`` (P,(max_sh_degree+1)^2,3) evaluated in-kernel with active degree ``d``.
    """
    g = torch.Generator().manual_seed(int(seed))
    cam = make_camera(W, H)
    U = lambda *s: torch.rand(*s, generator=g)
    Nrm = lambda *s: torch.randn(*s, generator=g)

    z = 2.0 + 4.0 * U(P)
    u, v = 2 * U(P) - 1, 2 * U(P) - 1
    redraw = U(P) < 0.02                      # 2 % behind / near the camera: exercises the near cull
    z = torch.where(redraw, -1.0 + 1.2 * U(P), z)
    x = u * z * cam["tanfovx"] * 1.1          # 10 % beyond the frustum: rect clamping, +-1.3 clamp
    y = v * z * cam["tanfovy"] * 1.1
    means3D = torch.stack([x, y, z], dim=1).float().contiguous()

    lo, hi = math.log(scale_range[0]), math.log(scale_range[1])
    s = torch.exp(lo + (hi - lo) * U(P))
    scales = (s[:, None] * torch.exp(0.3 * Nrm(P, 3))).float().contiguous()
    rot = Nrm(P, 4)
    if normalize_rot:
        rot = rot / rot.norm(dim=1, keepdim=True)
    rotations = rot.float().contiguous()
    opacities = (0.02 + 0.98 * U(P, 1)).float().contiguous()

    scene = dict(cam)
    scene.update(image_width=int(W), image_height=int(H), kernel_size=float(kernel_size), scale_modifier=1.0,
                 prefiltered=False, debug=False, means3D=means3D, opacities=opacities,
                 bg=torch.tensor(bg, dtype=torch.float32))
    if cov3D_precomp:
        # symmetric PSD 3x3 from scale/rot, upper triangle
        q = rotations / rotations.norm(dim=1, keepdim=True)
        r, xq, yq, zq = q.unbind(1)
        Rm = torch.stack([
            1 - 2 * (yq * yq + zq * zq), 2 * (xq * yq - r * zq), 2 * (xq * zq + r * yq),
            2 * (xq * yq + r * zq), 1 - 2 * (xq * xq + zq * zq), 2 * (yq * zq - r * xq),
            2 * (xq * zq - r * yq), 2 * (yq * zq + r * xq), 1 - 2 * (xq * xq + yq * yq)], dim=1).view(P, 3, 3)
        Mm = Rm * scales[:, None, :]
        Sig = Mm @ Mm.transpose(1, 2)
        scene["cov3D_precomp"] = torch.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2],
                                              Sig[:, 2, 2]], dim=1).float().contiguous()
    else:
        scene["scales"] = scales
        scene["rotations"] = rotations
    if sh_degree is None:
        scene["colors_precomp"] = U(P, 3).float().contiguous()
        scene["sh_degree"] = 0
    else:
        M = (max_sh_degree + 1) ** 2
        sh = 0.05 * Nrm(P, M, 3)
        sh[:, 0, :] = (U(P, 3) - 0.5) / C0
        scene["shs"] = sh.float().contiguous()
        scene["sh_degree"] = int(sh_degree)
    if subpixel_jitter > 0:
        scene["subpixel_offset"] = ((U(H, W, 2) - 0.5) * 2 * subpixel_jitter).float().contiguous()
    else:
        scene["subpixel_offset"] = torch.zeros((H, W, 2), dtype=torch.float32)
    g2 = torch.Generator().manual_seed(int(seed) + 1000)
    scene["dL_dpix"] = torch.randn(3, H, W, generator=g2).float().contiguous()
    return scene


CONFIGS = {
    # BASELINE.json configs
    "C1": dict(P=10_000, W=256, H=256, sh_degree=0, max_sh_degree=0),
    "C2": dict(P=500_000, W=800, H=800, sh_degree=3),
    "C3": dict(P=3_000_000, W=1920, H=1080, sh_degree=None),
    "C5": dict(P=6_000_000, W=4096, H=2160, sh_degree=None),
}


def to_device(scene, device):
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in scene.items()}
