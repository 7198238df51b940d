"""Opt-in fast path for the CALLER of the rasterizer, ``GaussianModel._render_internal`` (wildgaussians/method.py:1479-1632),
and for the statistics that consume its outputs.  ``method.py`` itself is not modified:

    import wildgaussians_fused
    wildgaussians_fused.enable(model)        # model: wildgaussians.method.GaussianModel

rebinds ``model._render_internal`` to ``render_internal`` below, which returns the same dictionary but

* evaluates the parameter activations + 3D filter (``get_gaussians``, method.py:1060-1086) with one kernel per direction;
* evaluates the raw and the appearance-toned colours with ONE fused kernel per direction (``fused_colors``:
  tcgen05 MLP + SH evaluation, ``csrc/appearance.cu``) instead of ~60 PyTorch launches and several P x 128 / P x 48
  fp32 intermediates (method.py:1570-1598);
* caches the per-camera constants -- the numpy 4x4 inversions, ``getWorld2View2``, the projection matrix, the camera
  centre, tan(fov) (method.py:1502-1527) -- per (pose, intrinsics, size), and the all-zero ``subpixel_offset`` /
  background tensors per (H, W, device): no per-call H2D copies, ``zeros(H, W, 2)`` fills or device-side 4x4 inverse;
* does not synchronise on ``active_sh_degree.cpu().item()`` (method.py:1540) every call: the value is read back only when
  the buffer's version counter changed;
* both rasterizer passes share one projection / depth order / tile binning (the drop-in package's geometry reuse).

``add_densification_stats`` (model method) / ``densification_stats`` (function) fuse method.py:1997-1998 and :1470-1477
into one kernel.  Only the reference's default configuration is accelerated (appearance on with separate tuned colour,
``appearance_model_sh`` off, SH degree 3, 24 + 32 embedding features); anything else falls through to the original
method, so ``enable`` is always safe.
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np
import torch

from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer, _C
from fused_colors import fused_activations, fused_colors

_cam_cache: dict = {}
_zero_cache: dict = {}
_deg_cache: dict = {}
_CAM_CACHE_MAX = 4096


def _camera_constants(m, cam, device):
    """The tensors / scalars of method.py:1502-1527 for one camera, computed once with the reference's own helper functions."""
    poses = np.ascontiguousarray(cam.poses)
    intr = np.asarray(cam.intrinsics, dtype=np.float64)
    width, height = (int(x) for x in cam.image_sizes)
    key = (poses.tobytes(), intr.tobytes(), width, height, str(device))
    hit = _cam_cache.get(key)
    if hit is not None:
        return hit
    pose = np.concatenate([np.copy(poses), np.array([[0, 0, 0, 1]], dtype=poses.dtype)], axis=0)
    pose = np.linalg.inv(pose)
    R = np.transpose(pose[:3, :3])
    T = pose[:3, 3]
    fx, fy, cx, cy = cam.intrinsics
    world_view = torch.tensor(m.getWorld2View2(R, T, np.array([0.0, 0.0, 0.0], dtype=np.float32), 1.0)).transpose(0, 1).to(device=device)
    # (float() arguments: the helper assigns its results into a torch tensor element, which rejects numpy.float32 scalars)
    proj = m.getProjectionMatrixFromOpenCV(width, height, float(fx), float(fy), float(cx), float(cy), 0.01, 100.0) \
        .transpose(0, 1).to(device=device)
    full_proj = (world_view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    cam_center = world_view.inverse()[3, :3]
    out = dict(world_view=world_view.contiguous(), full_proj=full_proj.contiguous(), cam_center=cam_center.contiguous(),
               tanfovx=math.tan(m.focal2fov(float(fx), float(width)) * 0.5),
               tanfovy=math.tan(m.focal2fov(float(fy), float(height)) * 0.5), width=width, height=height)
    if len(_cam_cache) >= _CAM_CACHE_MAX:
        _cam_cache.clear()
    _cam_cache[key] = out
    return out


def _zeros(shape, device):
    key = (tuple(shape), str(device))
    t = _zero_cache.get(key)
    if t is None:
        t = _zero_cache[key] = torch.zeros(shape, dtype=torch.float32, device=device)
    return t


def _active_degree(model) -> int:
    t = model.active_sh_degree
    key = id(t)
    hit = _deg_cache.get(key)
    if hit is None or hit[0] != t._version or hit[2] is not t:
        hit = _deg_cache[key] = (t._version, int(t.item()), t)
    return hit[1]


def _supported(model, config) -> bool:
    c = model.config
    return (bool(c.appearance_enabled) and bool(c.appearance_separate_tuned_color) and not bool(c.appearance_model_sh)
            and int(c.sh_degree) == 3 and model.features_rest is not None and model.embeddings is not None
            and model.appearance_mlp is not None and model.embeddings.shape[1] == 24 and int(c.appearance_embedding_dim) == 32
            and model.xyz.is_cuda and not bool(config.debug))


def render_internal(self, viewpoint_camera, config, *, kernel_size, scaling_modifier=1.0, embedding, return_raw=True,
                    render_depth=False):
    """Drop-in for ``GaussianModel._render_internal``: same arguments, same output dictionary."""
    if embedding is None or not _supported(self, config):
        return type(self)._render_internal(self, viewpoint_camera, config, kernel_size=kernel_size,
                                           scaling_modifier=scaling_modifier, embedding=embedding, return_raw=return_raw,
                                           render_depth=render_depth)
    m = sys.modules[type(self).__module__]
    device = self.xyz.device
    assert len(viewpoint_camera.poses.shape) == 2, "Expected a single camera"
    assert viewpoint_camera.image_sizes is not None, "Expected image sizes to be set"
    cc = _camera_constants(m, viewpoint_camera, device)
    H, W = cc["height"], cc["width"]

    # zero tensor whose .grad receives the screen-space gradients (method.py:1495-1499)
    screenspace_points = torch.zeros_like(self.xyz, dtype=self.xyz.dtype, requires_grad=True, device=device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    deg = _active_degree(self)
    settings = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cc["tanfovx"], tanfovy=cc["tanfovy"], kernel_size=kernel_size,
        subpixel_offset=_zeros((H, W, 2), device), bg=_zeros((3,), device), scale_modifier=scaling_modifier,
        viewmatrix=cc["world_view"], projmatrix=cc["full_proj"], sh_degree=deg, campos=cc["cam_center"], prefiltered=False,
        debug=False, return_accumulation=True)
    rasterizer = GaussianRasterizer(raster_settings=settings)

    # activations + 3D filter (get_gaussians, method.py:1060-1086) in one kernel per direction
    means3D = self.xyz
    scales, opacity, rotations = fused_activations(self.scales, self.opacities, self.rotations, self.filter_3D)
    colors_raw, colors_toned = fused_colors(self.features_dc, self.features_rest, self.embeddings, embedding,
                                            self.appearance_mlp.mlp, means3D, cc["cam_center"], deg, want_raw=bool(return_raw))
    kw = dict(means3D=means3D, means2D=screenspace_points, opacities=opacity, scales=scales, rotations=rotations, shs=None,
              cov3D_precomp=None)
    raw_image = radii = accumulation = None
    if return_raw:
        raw_image, radii, accumulation = rasterizer(colors_precomp=colors_raw, **kw)
    image, radii2, accumulation2 = rasterizer(colors_precomp=colors_toned, **kw)
    radii = radii2 if radii is None else radii
    accumulation = accumulation2 if accumulation is None else accumulation
    out = {"render": image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
           "accumulation": accumulation, "radii": radii}
    if return_raw:
        out["raw_render"] = raw_image
    if render_depth:
        dist = torch.norm(means3D - cc["cam_center"][None], dim=-1).unsqueeze(-1).repeat(1, 3)
        out["depth"] = rasterizer(colors_precomp=dist, **kw)[0][0]
    return out


def densification_stats(model, viewspace_point_tensor, radii) -> None:
    """method.py:1997-1998 + ``add_densification_stats`` (:1470-1477) in one kernel (in place on the model's buffers)."""
    grad = viewspace_point_tensor.grad
    P = int(radii.shape[0])
    dev = radii.device
    use_abs = bool(model.config.use_gof_abs_gradient)
    bufs = [model.max_radii2D, model.xyz_grad, model.denom] + ([model.xyz_gradient_accum_abs, model.xyz_gradient_accum_abs_max] if use_abs else [])
    if not (radii.is_cuda and grad is not None and grad.dtype == torch.float32 and grad.is_contiguous() and radii.dtype == torch.int32
            and radii.is_contiguous() and all(b.dtype == torch.float32 and b.is_contiguous() and b.numel() == P for b in bufs)):
        vis = radii > 0
        model.max_radii2D[vis] = torch.max(model.max_radii2D[vis], radii[vis])
        model.add_densification_stats(viewspace_point_tensor, vis)
        return
    with torch.cuda.device(dev):
        _C._check(_C._lib.gsr_densification_stats(
            P, radii.data_ptr(), grad.data_ptr(), model.max_radii2D.data_ptr(), model.xyz_grad.data_ptr(),
            model.xyz_gradient_accum_abs.data_ptr() if use_abs else None,
            model.xyz_gradient_accum_abs_max.data_ptr() if use_abs else None, model.denom.data_ptr(),
            torch.cuda.current_stream(dev).cuda_stream), "gsr_densification_stats")


def camera_table(cameras) -> tuple:
    """The per-camera constants of ``compute_3D_filter`` (method.py:1147-1181) for all cameras as one [C, 20] float32 array
    (layout: include/gsrast.h, gsr_compute_3d_filter) and the focal length of the highest-resolution camera (:1178-1179).
    Built with the reference's own statements (numpy 4x4 inverse, transposed rotation) in one pass on the host."""
    rows, focal = [], 0.0
    for camera in cameras:
        assert camera.image_sizes is not None, "Camera image size is not set"
        fx, fy, _, _ = camera.intrinsics
        width, height = camera.image_sizes
        pose = np.copy(camera.poses)
        pose = np.concatenate([pose, np.array([[0, 0, 0, 1]], dtype=pose.dtype)], axis=0)
        pose = np.linalg.inv(pose)
        R = np.transpose(pose[:3, :3]).astype(np.float32)
        T = pose[:3, 3].astype(np.float32)
        f32 = np.float32
        rows.append(np.concatenate([R.reshape(-1), T, np.array(
            [f32(fx), f32(fy), f32(width / 2.0), f32(height / 2.0), f32(-0.15 * width), f32(width * 1.15), f32(-0.15 * height),
             f32(1.15 * height)], dtype=np.float32)]))
        if focal < fx:
            focal = fx
    table = np.stack(rows).astype(np.float32) if rows else np.zeros((0, 20), dtype=np.float32)
    return table, float(focal)


def compute_3D_filter(model, cameras) -> None:
    """``GaussianModel.compute_3D_filter`` (method.py:1140-1190) in two kernels over all cameras (csrc/filter3d.cu) instead of
    ~20 PyTorch launches + 2 H2D copies + 2 host synchronisations PER CAMERA; registers the same ``filter_3D`` buffer."""
    xyz = model.xyz
    if not (xyz.is_cuda and xyz.dtype == torch.float32 and xyz.is_contiguous()):
        return type(model).compute_3D_filter(model, cameras)
    table, focal = camera_table(cameras)
    if table.shape[0] == 0 or not focal > 0:
        return type(model).compute_3D_filter(model, cameras)
    dev, P = xyz.device, int(xyz.shape[0])
    with torch.no_grad(), torch.cuda.device(dev):
        cams = torch.from_numpy(table).to(dev)
        scratch = torch.empty((_C._lib.gsr_filter3d_scratch_bytes(P) + 15) // 16 * 4, dtype=torch.float32, device=dev)
        out = torch.empty((P,), dtype=torch.float32, device=dev)
        _C._check(_C._lib.gsr_compute_3d_filter(P, xyz.data_ptr(), int(table.shape[0]), cams.data_ptr(), focal, out.data_ptr(),
                                                scratch.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                  "gsr_compute_3d_filter")
        filter_3D = out.to(dtype=model.filter_3D.dtype, device=model.filter_3D.device)
        del model.filter_3D
        model.register_buffer("filter_3D", filter_3D[..., None])


def enable(model, optimizer: bool = True):
    """Route this model's ``_render_internal`` and ``compute_3D_filter`` through the fused paths (instance attributes; the
    class is untouched) and, with ``optimizer``, turn its Adam into the single-kernel ``fused_adam.FusedAdam`` in place."""
    model._render_internal = types.MethodType(render_internal, model)
    model.compute_3D_filter = types.MethodType(compute_3D_filter, model)
    if optimizer and type(getattr(model, "optimizer", None)) is torch.optim.Adam:
        import fused_adam
        fused_adam.adopt(model.optimizer)
    return model


def disable(model):
    model.__dict__.pop("_render_internal", None)
    model.__dict__.pop("compute_3D_filter", None)
    if type(getattr(model, "optimizer", None)).__name__ == "FusedAdam":
        model.optimizer.__class__ = torch.optim.Adam
    return model
