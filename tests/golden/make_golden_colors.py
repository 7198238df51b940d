"""Golden vectors of wild-gaussians' per-Gaussian colour path, produced by the REFERENCE ITSELF.

Runs in the build container only (needs /root/reference; CPU torch is enough):
    python tests/golden/make_golden_colors.py
imports wildgaussians/method.py unmodified -- its unavailable third-party imports (omegaconf, plyfile, simple_knn),
none of which is touched on this path, are stubbed, and `diff_gaussian_rasterization` resolves to this repo's drop-in
package -- builds `EmbeddingModel(Config(...))` under a seed and records inputs, weights, outputs and autograd
gradients of exactly the statements `_render_internal` executes for the toned colours (method.py:1586-1598).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def import_reference():
    for p in (os.path.join(ROOT, "wild-gaussians_b200"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:                                        # this repo's drop-in for the reference's simple-knn submodule
        import simple_knn._C  # noqa: F401
    except Exception:
        pass
    for name in ("omegaconf", "plyfile", "simple_knn", "simple_knn._C"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["omegaconf"].OmegaConf = object
    sys.modules["plyfile"].PlyData = object
    sys.modules["plyfile"].PlyElement = object
    if not hasattr(sys.modules["simple_knn._C"], "distCUDA2"):
        sys.modules["simple_knn._C"].distCUDA2 = None
    import wildgaussians.method as m
    from wildgaussians.config import Config
    return m, Config


def make_case(m, Config, P, active_deg, seed, trained_like=False):
    torch.manual_seed(seed)
    cfg = Config(source_path="", model_path="")
    model = m.EmbeddingModel(cfg).double()
    if trained_like:
        # at initialisation `mul` (= 0.01 x MLP output) is ~0 and every toned colour is ~0.5; a trained model has
        # mul ~ 1: emulate it through the last layer's bias so that the clamps at 0 and 1 are exercised
        with torch.no_grad():
            model.mlp[-1].bias[3:] = 100.0
    g = torch.Generator().manual_seed(seed + 1)
    feats = torch.randn(P, 48, generator=g, dtype=torch.float64) * 0.3
    feats[:, :3] = (torch.rand(P, 3, generator=g, dtype=torch.float64) * 1.6 - 0.9) / 0.28209479177387814   # some colours clamp at 0
    gemb = torch.randn(P, 24, generator=g, dtype=torch.float64) * 0.5
    aemb = torch.randn(32, generator=g, dtype=torch.float64) * 0.5
    means3D = torch.randn(P, 3, generator=g, dtype=torch.float64) * 2.0
    campos = torch.tensor([0.1, -0.2, 0.3], dtype=torch.float64)
    dL = torch.randn(P, 3, generator=g, dtype=torch.float64)
    feats.requires_grad_(True); gemb.requires_grad_(True); aemb.requires_grad_(True)
    # ---- the statements of _render_internal (method.py:1570,1586-1598), verbatim in structure
    features = feats.clamp_max(1.0)
    dir_pp_normalized = torch.nn.functional.normalize(means3D - camera_center_repeat(campos, P), dim=1)
    embedding_expanded = aemb[None].repeat(P, 1)
    colors_toned = model(gemb, embedding_expanded, features).clamp_max(1.0)
    shdim = (cfg.sh_degree + 1) ** 2
    colors_toned = colors_toned.view(-1, shdim, 3).transpose(1, 2).contiguous().clamp_max(1.0)
    colors_toned = m.eval_sh(active_deg, colors_toned, dir_pp_normalized)
    colors_toned = torch.clamp_min(colors_toned + 0.5, 0.0)
    (colors_toned * dL).sum().backward()
    lin = [l for l in model.mlp if isinstance(l, torch.nn.Linear)]
    out = dict(P=P, active_deg=active_deg, features=feats.detach().numpy(), gembedding=gemb.detach().numpy(),
               aembedding=aemb.detach().numpy(), means3D=means3D.numpy(), campos=campos.numpy(), dL_dcolors=dL.numpy(),
               colors=colors_toned.detach().numpy(), g_features=feats.grad.numpy(), g_gembedding=gemb.grad.numpy(),
               g_aembedding=aemb.grad.numpy())
    for i, l in enumerate(lin):
        out[f"W{i + 1}"] = l.weight.detach().numpy(); out[f"b{i + 1}"] = l.bias.detach().numpy()
        out[f"g_W{i + 1}"] = l.weight.grad.numpy(); out[f"g_b{i + 1}"] = l.bias.grad.numpy()
    return out


def camera_center_repeat(campos, P):
    return campos.repeat(P, 1)


if __name__ == "__main__":
    m, Config = import_reference()
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (P, deg, seed, trained) in {"colors_deg3": (400, 3, 11, True), "colors_deg1": (200, 1, 12, False)}.items():
        case = make_case(m, Config, P, deg, seed, trained)
        np.savez_compressed(os.path.join(here, name + ".npz"), **case)
        print(name, "colors", case["colors"].shape, "clamped-at-0:", int((case["colors"] == 0).sum()))
