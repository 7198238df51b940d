"""Helpers shared by the golden-vector tests."""
import os

import numpy as np

import synthetic
from make_golden import CASES, input_digest

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRAD_NAMES = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales",
              "dL_drotations")
# tolerances of BASELINE.json's north star
PIX_TOL = 1e-4      # absolute, on pixels
GRAD_TOL = 1e-3     # relative to the largest |gradient| of the tensor


def load_case(name):
    """(scene, golden dict).  The scene's scalars come from the generator; its TENSORS are the committed inputs the golden
    vectors were made from (tests/golden/<name>.inputs.npz): regenerating them from the seed is not bit-reproducible across
    host CPUs (torch's vectorised randn / exp / log differ in the last bit between AVX2 and AVX-512 hosts -- observed on
    one GPU box for c1_deg0), and the goldens are compared bit-exactly.  The stored digest pins the pair."""
    import torch
    scene = synthetic.make_scene(**CASES[name])
    gold = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    inputs = os.path.join(GOLDEN_DIR, name + ".inputs.npz")
    if os.path.exists(inputs):
        for k, v in np.load(inputs).items():
            assert k in scene and tuple(scene[k].shape) == v.shape, k
            scene[k] = torch.from_numpy(np.ascontiguousarray(v))
    assert str(gold["input_digest"]) == input_digest(scene), \
        "the inputs no longer match the digest the golden vectors were made from"
    return scene, gold


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def rel_err(a, b):
    scale = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max()) / scale
