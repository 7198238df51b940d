"""CPU tests of the host-side logic added around the hot path: the per-camera constant cache of the fused caller
 against the statements of the reference it replaces, the band sentinel of the C ABI, the geometry-cache
key, and the one-shot composite-input deferral."""
import math
import os
import sys
import types

import numpy as np
import pytest
import torch

import wg_harness as wh


def test_camera_constants_match_the_reference_statements():
    """wildgaussians_fused._camera_constants == method.py:1502-1527, bit for bit, and is cached per camera."""
    m, Config = wh.import_method()
    if m is None:
        pytest.skip("reference python package not present")
    import wildgaussians_fused as wf
    rng = np.random.default_rng(3)
    # a rotated / translated camera-to-world pose, non-centred principal point
    A = rng.normal(size=(3, 3)); Q, _ = np.linalg.qr(A)
    pose = np.concatenate([Q, rng.normal(size=(3, 1))], axis=1).astype(np.float32)
    width, height = 1237, 811
    intr = np.array([1000.5, 990.25, 600.0, 400.5], dtype=np.float32)
    cam = types.SimpleNamespace(poses=pose, image_sizes=np.array([width, height], dtype=np.int32), intrinsics=intr)
    dev = torch.device("cpu")
    cc = wf._camera_constants(m, cam, dev)
    assert wf._camera_constants(m, cam, dev) is cc                      # cached
    # ---- the reference's statements (method.py:1502-1527)
    p = np.copy(cam.poses)
    p = np.concatenate([p, np.array([[0, 0, 0, 1]], dtype=p.dtype)], axis=0)
    p = np.linalg.inv(p)
    R = np.transpose(p[:3, :3]); T = p[:3, 3]
    fx, fy, cx, cy = cam.intrinsics
    wv = torch.tensor(m.getWorld2View2(R, T, np.array([0.0, 0.0, 0.0], dtype=np.float32), 1.0)).transpose(0, 1)
    proj = m.getProjectionMatrixFromOpenCV(width, height, float(fx), float(fy), float(cx), float(cy), 0.01, 100.0).transpose(0, 1)
    full = (wv.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    center = wv.inverse()[3, :3]
    assert torch.equal(cc["world_view"], wv.contiguous()) and cc["world_view"].is_contiguous()
    assert torch.equal(cc["full_proj"], full) and torch.equal(cc["cam_center"], center)
    assert cc["tanfovx"] == math.tan(m.focal2fov(float(fx), float(width)) * 0.5)
    assert cc["tanfovy"] == math.tan(m.focal2fov(float(fy), float(height)) * 0.5)
    assert (cc["width"], cc["height"]) == (width, height)
    # another pose -> another entry
    cam2 = types.SimpleNamespace(poses=pose + 0.01, image_sizes=cam.image_sizes, intrinsics=intr)
    assert wf._camera_constants(m, cam2, dev) is not cc


def test_active_degree_is_read_back_only_when_the_buffer_changes():
    import wildgaussians_fused as wf
    model = types.SimpleNamespace(active_sh_degree=torch.full((), 1, dtype=torch.int32))
    assert wf._active_degree(model) == 1
    calls = []
    orig = torch.Tensor.item
    try:
        torch.Tensor.item = lambda self: (calls.append(1), orig(self))[1]
        assert wf._active_degree(model) == 1 and not calls               # cached: no read-back
        model.active_sh_degree += 1                                      # in-place update bumps the version
        assert wf._active_degree(model) == 2 and len(calls) == 1
    finally:
        torch.Tensor.item = orig


def test_band_sentinel_and_geometry_key():
    from diff_gaussian_rasterization import _C
    assert _C._abi_shard((0, 0), 1080) == (0, 0)                         # whole image
    assert _C._abi_shard((3, 7), 1080) == (3, 7)
    assert _C._abi_shard((5, 5), 1080) == (68, 68)                       # an EMPTY band is not the whole-image sentinel
    assert _C._abi_shard((0, 0 + 0), 100) == (0, 0)
    a, b = torch.zeros(4, 3), torch.zeros(4, 3)
    e1, e2 = torch.Tensor([]), torch.Tensor([])
    k1 = _C._geometry_key("cpu", 0, 4, 8, 8, 0, (1.0,), (0, 0), (a, e1))
    assert k1 == _C._geometry_key("cpu", 0, 4, 8, 8, 0, (1.0,), (0, 0), (a, e2))     # absent inputs compare equal
    assert k1 != _C._geometry_key("cpu", 0, 4, 8, 8, 0, (1.0,), (0, 0), (b, e1))     # another tensor object
    a.add_(1)
    assert k1 != _C._geometry_key("cpu", 0, 4, 8, 8, 0, (1.0,), (0, 0), (a, e1))     # in-place update bumps the version
    assert k1 != _C._geometry_key("cpu", 0, 4, 8, 8, 0, (1.0,), (2, 5), (a, e1))


def test_fused_entry_points_reject_cpu_tensors_without_a_gpu():
    import fused_colors as fc
    mlp = torch.nn.Sequential(torch.nn.Linear(59, 128), torch.nn.ReLU(), torch.nn.Linear(128, 128), torch.nn.ReLU(),
                              torch.nn.Linear(128, 6))
    with pytest.raises(RuntimeError, match="CUDA"):
        fc.fused_colors(torch.zeros(4, 3), torch.zeros(4, 45), torch.zeros(4, 24), torch.zeros(32), mlp, torch.zeros(4, 3),
                        torch.zeros(3), 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        fc.fused_activations(torch.zeros(4, 3), torch.zeros(4, 1), torch.zeros(4, 4), torch.zeros(4, 1))


def test_reduction_mode_default_depends_on_world_size(monkeypatch):
    """parallel.peer_mode(): peer `red` for two ranks, the pull form from three ranks on; GSR_PEER_REDUCE overrides."""
    import parallel
    import torch.distributed as dist
    monkeypatch.setattr(parallel, "_PEER_MODE_ENV", None)
    for world, want in ((2, 1), (3, 3), (4, 3), (8, 3)):
        monkeypatch.setattr(dist, "get_world_size", lambda group=None, w=world: w)
        assert parallel.peer_mode() == want
    monkeypatch.setattr(parallel, "_PEER_MODE_ENV", "0")
    assert parallel.peer_mode() == 0


def test_pull_mode_entry_points_reject_bad_arguments():
    """gsr_backward_partials_marked / gsr_backward_finalize_pull validate their arguments before touching CUDA."""
    import ctypes
    import diff_gaussian_rasterization._C as C
    a = C.GsrBackwardArgs()
    a.P = 10
    assert C._lib.gsr_backward_partials_marked(ctypes.byref(a), None, None) != 0            # touched is NULL
    assert b"touched" in C._lib.gsr_last_error()
    arr = (ctypes.c_void_p * 9)()
    a.P, a.W, a.H = 0, 16, 16
    assert C._lib.gsr_backward_finalize_pull(ctypes.byref(a), arr, arr, 2, 0, None, None, None) == 0     # P == 0: nothing to do
