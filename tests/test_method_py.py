"""The reference's own ``wildgaussians/method.py``, UNMODIFIED, running on the drop-in package (SURVEY.md 4.5,
BASELINE.json north star: "so wildgaussians/method.py drops onto it unchanged").

``GaussianModel._render_internal`` (method.py:1479-1632; appearance on, uncertainty off) is executed once on this repo's
``diff_gaussian_rasterization`` and once on the reference's compiled CUDA rasterizer behind the same two names; images,
radii, accumulation and every parameter gradient of a train-step-like loss must agree.  It is also the real call
pattern of the geometry reuse: two rasterizer calls per step on the same geometry tensors, the view
matrix a non-contiguous transposed tensor (method.py:1516) -> exactly one cache hit per step.
"""
import os
import sys

import pytest
import torch

import synthetic
import wg_harness as wh

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref_api():
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("oracle/_ref/libdgr_ref.so not present")
    sys.path.insert(0, ROOT)
    import bench
    return bench.make_reference_api(ref_cuda)


@pytest.mark.parametrize("kw", [dict(P=150_000, W=640, H=400, seed=61), dict(P=40_000, W=333, H=211, seed=62)],
                         ids=lambda k: f"P{k['P']}_{k['W']}x{k['H']}")
def test_render_internal_on_the_drop_in_equals_reference_backend(kw, ref_api):
    m, Config = wh.import_method()
    if m is None:
        pytest.skip("reference python package not present (baseline/_ref)")
    import diff_gaussian_rasterization as ours
    from diff_gaussian_rasterization import _C
    dev = torch.device("cuda:0")
    scene = synthetic.make_scene(sh_degree=3, **kw)
    model, cfg = wh.make_model(m, Config, scene, dev, seed=kw["seed"])
    cam = wh.make_camera(scene)
    g = torch.Generator().manual_seed(kw["seed"] + 100)
    G1 = torch.randn(3, kw["H"], kw["W"], generator=g).to(dev)
    G2 = torch.randn(3, kw["H"], kw["W"], generator=g).to(dev)

    # ---- ours: the names method.py imported at :26 ARE this package's
    wh.use_backend(m, ours.GaussianRasterizer, ours.GaussianRasterizationSettings)
    _C.set_geometry_cache(True)
    _C.clear_geometry_cache()
    h0 = _C.geometry_cache_hits()
    a = wh.collect(model, wh.train_step(model, cfg, cam, G1, G2))
    assert _C.geometry_cache_hits() == h0 + 1, "the second composite of the step did not reuse the first one's geometry"
    a2 = wh.collect(model, wh.train_step(model, cfg, cam, G1, G2))      # next step: fresh tensors -> miss, then hit
    assert _C.geometry_cache_hits() == h0 + 2
    assert torch.equal(a["render"], a2["render"]) and torch.equal(a["raw_render"], a2["raw_render"])

    # ---- the reference's compiled CUDA core behind the same two names
    wh.use_backend(m, ref_api["GaussianRasterizer"], ref_api["GaussianRasterizationSettings"])
    b = wh.collect(model, wh.train_step(model, cfg, cam, G1, G2))
    b2 = wh.collect(model, wh.train_step(model, cfg, cam, G1, G2))      # the reference's own atomic-order noise
    wh.use_backend(m, ours.GaussianRasterizer, ours.GaussianRasterizationSettings)
    torch.cuda.synchronize()

    assert torch.equal(a["radii"], b["radii"])
    assert float((a["render"] - b["render"]).abs().max()) <= 1e-4
    assert float((a["raw_render"] - b["raw_render"]).abs().max()) <= 1e-4
    assert float((a["accumulation"] - b["accumulation"]).abs().max()) <= 1e-4
    assert float(a["render"].abs().max()) > 0.05 and float((a["render"] - a["raw_render"]).abs().max()) > 1e-3
    for k in a:
        if not (k.startswith("g_") or k == "viewspace_grad"):
            continue
        scale = float(b[k].abs().max()) + 1e-30
        err = float((a[k] - b[k]).abs().max()) / scale
        noise = float((b2[k] - b[k]).abs().max()) / scale
        assert err <= max(20.0 * noise, 2e-5), (k, err, noise)
        assert err < 1e-3, (k, err)
