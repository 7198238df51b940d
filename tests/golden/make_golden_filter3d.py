"""Generates tests/golden/filter3d_*.npz by running the reference's OWN, unmodified ``GaussianModel.compute_3D_filter``
(wildgaussians/method.py:1140-1190, imported from /root/reference) on the CPU on a seeded cloud and seeded cameras.
Run in the build container (needs /root/reference): python tests/golden/make_golden_filter3d.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "wild-gaussians_b200"), ROOT]


def make_cameras(C, seed, scale=4.0):
    """C pinhole cameras on a rough sphere of radius ~`scale` around the origin looking inwards (camera-to-world 3x4)."""
    rng = np.random.default_rng(seed)
    cams = []
    for i in range(C):
        eye = rng.normal(size=3); eye = eye / np.linalg.norm(eye) * scale * rng.uniform(0.6, 1.4)
        fwd = -eye / np.linalg.norm(eye) + rng.normal(size=3) * 0.15
        fwd /= np.linalg.norm(fwd)
        up = np.array([0.0, 1.0, 0.0]) if abs(fwd[1]) < 0.9 else np.array([1.0, 0.0, 0.0])
        right = np.cross(up, fwd); right /= np.linalg.norm(right)
        up2 = np.cross(fwd, right)
        c2w = np.stack([right, up2, fwd, eye], axis=1).astype(np.float32)          # columns: x, y, z axes, position
        W, H = int(rng.integers(300, 900)), int(rng.integers(200, 700))
        f = float(rng.uniform(0.7, 1.4) * W)
        cams.append(types.SimpleNamespace(poses=c2w, image_sizes=np.array([W, H], dtype=np.int32),
                                          intrinsics=np.array([f, f * rng.uniform(0.95, 1.05), W / 2.0, H / 2.0], dtype=np.float32)))
    return cams


def main():
    import wg_harness as wh
    m, Config = wh.import_method()
    assert m is not None
    for name, P, C, seed in (("small", 3000, 7, 1), ("many_cams", 1500, 300, 2)):
        cfg = Config(source_path="", model_path="", uncertainty_mode="disabled")
        model = m.GaussianModel(cfg, training_setup=False)
        model._resize_parameters(P)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            model.xyz.copy_(torch.randn(P, 3, generator=g) * 2.5)
        cams = make_cameras(C, seed + 10)
        model.compute_3D_filter(cams)
        np.savez_compressed(os.path.join(HERE, f"filter3d_{name}.npz"), xyz=model.xyz.detach().numpy(),
                            poses=np.stack([c.poses for c in cams]), sizes=np.stack([c.image_sizes for c in cams]),
                            intrinsics=np.stack([c.intrinsics for c in cams]), filter_3D=model.filter_3D.numpy())
        print(name, model.filter_3D.shape, float(model.filter_3D.min()), float(model.filter_3D.max()))


if __name__ == "__main__":
    main()
