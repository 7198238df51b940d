"""TEST INFRASTRUCTURE -- CPU restatement (numpy, fp64 or fp32) of wild-gaussians' per-Gaussian colour path, the
"next" row SURVEY.md 8f-2 (not built yet: ROADMAP.md 1).  Only tests/ may import this module.

What it restates (paths relative to /root/reference/wildgaussians/):
  * EmbeddingModel.forward              method.py:874-900   59 -> 128 -> 128 -> 6 MLP, x 0.01, affine on the features
  * the toned-colour evaluation         method.py:1589-1598 view(-1, 16, 3) -> eval_sh -> + 0.5 -> clamp_min(0)
  * eval_sh                             method.py:493-548   (degrees 0..3; same polynomial as DGR forward.cu:20-71)

Pinned by tests/golden/colors_*.npz, which were produced by importing the reference itself in the build container
(tests/golden/make_golden_colors.py) -- see tests/test_color_oracle.py.
"""
from __future__ import annotations

import numpy as np

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435)


def eval_sh(deg: int, sh: np.ndarray, dirs: np.ndarray) -> np.ndarray:
    """sh: [P, C, K] coefficients, dirs: [P, 3] unit directions -> [P, C]   (method.py:493-548)"""
    assert 0 <= deg <= 3 and sh.shape[-1] >= (deg + 1) ** 2
    res = C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
        res = res - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
                   + C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                res = (res + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10]
                       + C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                       + C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14]
                       + C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return res


def appearance_features(features, gembedding, aembedding, weights):
    """EmbeddingModel.forward with appearance_model_sh=False (method.py:889-900).
    features [P, 48] (DC first), gembedding [P, 24], aembedding [P or 1, 32]; weights = (W1[128,59], b1, W2[128,128],
    b2, W3[6,128], b3) in torch.nn.Linear layout.  Returns the toned features [P, 48]."""
    W1, b1, W2, b2, W3, b3 = weights
    P = features.shape[0]
    if aembedding.shape[0] == 1:
        aembedding = np.repeat(aembedding, P, axis=0)
    inp = np.concatenate([features[:, :3], gembedding, aembedding], axis=1)
    h = np.maximum(inp @ W1.T + b1, 0.0)
    h = np.maximum(h @ W2.T + b2, 0.0)
    out = (h @ W3.T + b3) * 0.01
    offset, mul = out[:, :3], out[:, 3:]
    offset_full = np.concatenate([offset / C0, np.zeros((P, features.shape[1] - 3), dtype=features.dtype)], axis=1)
    mul_full = np.tile(mul, (1, features.shape[1] // 3))
    return features * mul_full + offset_full


def toned_colors(features, gembedding, aembedding, weights, means3D, campos, active_sh_degree, sh_degree=3):
    """colors_toned of GaussianModel._render_internal (method.py:1589-1598): appearance MLP -> clamp_max(1) ->
    view(-1, K, 3).transpose(1, 2) -> clamp_max(1) -> eval_sh(active degree) -> + 0.5 -> clamp_min(0)."""
    K = (sh_degree + 1) ** 2
    toned = np.minimum(appearance_features(features, gembedding, aembedding, weights), 1.0)
    sh = np.minimum(toned.reshape(-1, K, 3).transpose(0, 2, 1), 1.0)        # [P, 3, K]
    d = means3D - campos[None]
    dirs = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)  # F.normalize(dim=1)
    return np.maximum(eval_sh(active_sh_degree, sh, dirs) + 0.5, 0.0)
