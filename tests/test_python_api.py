"""CPU tests of the drop-in Python surface (names, field order, argument validation, loud failure without CUDA)."""
import inspect

import pytest
import torch

import diff_gaussian_rasterization as dgr
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer


def _settings(H=32, W=48):
    return GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=0.5, tanfovy=0.5, kernel_size=0.1,
        subpixel_offset=torch.zeros(H, W, 2), bg=torch.zeros(3), scale_modifier=1.0, viewmatrix=torch.eye(4),
        projmatrix=torch.eye(4), sh_degree=0, campos=torch.zeros(3), prefiltered=False, debug=False,
        return_accumulation=True)


def test_settings_fields_match_reference_order():
    # submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py:175-190
    assert GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "kernel_size", "subpixel_offset", "bg",
        "scale_modifier", "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered", "debug",
        "return_accumulation")


def test_public_names():
    assert hasattr(dgr, "_C") and hasattr(dgr, "rasterize_gaussians") and hasattr(dgr, "_RasterizeGaussians")
    for fn in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
        assert callable(getattr(dgr._C, fn))
    assert len(inspect.signature(dgr._C.rasterize_gaussians).parameters) == 21
    assert len(inspect.signature(dgr._C.rasterize_gaussians_backward).parameters) == 23
    assert len(inspect.signature(dgr._C.mark_visible).parameters) == 3
    r = GaussianRasterizer(_settings())
    assert isinstance(r, torch.nn.Module) and r.raster_settings.image_width == 48
    params = list(inspect.signature(r.forward).parameters)
    assert params == ["means3D", "means2D", "opacities", "shs", "colors_precomp", "scales", "rotations",
                      "cov3D_precomp"]


def test_exclusivity_checks_raise_like_the_reference():
    r = GaussianRasterizer(_settings())
    m = torch.zeros(4, 3); o = torch.ones(4, 1); s = torch.ones(4, 3); q = torch.ones(4, 4); c = torch.ones(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, m, o, scales=s, rotations=q)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, m, o, shs=torch.ones(4, 1, 3), colors_precomp=c, scales=s, rotations=q)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, o, colors_precomp=c)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, o, colors_precomp=c, scales=s, rotations=q, cov3D_precomp=torch.ones(4, 6))


def test_no_silent_cpu_fallback():
    r = GaussianRasterizer(_settings())
    m = torch.zeros(4, 3); o = torch.ones(4, 1); s = torch.ones(4, 3); q = torch.ones(4, 4); c = torch.ones(4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(m, m, o, colors_precomp=c, scales=s, rotations=q)
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        dgr._C.rasterize_gaussians(torch.zeros(3), torch.zeros(4), c, o, s, q, 1.0, torch.Tensor([]), torch.eye(4),
                                   torch.eye(4), 0.5, 0.5, 0.1, torch.zeros(32, 48, 2), 32, 48, torch.Tensor([]), 0,
                                   torch.zeros(3), False, False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        r.markVisible(m)


def test_product_does_not_import_the_oracle():
    import os, re
    pkg = os.path.dirname(os.path.dirname(os.path.abspath(dgr.__file__)))
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|cpu_oracle|ref_cuda|liboracle|libdgr_ref", src, re.M), f
