"""Generates tests/golden/knn_*.npz on a GPU box by running the UNMODIFIED reference simple-knn (oracle/_ref/libsimpleknn_ref.so,
built by `make -C oracle knn_ref` from /root/reference) on seeded point clouds:  python tests/golden/make_golden_knn.py
(inputs are regenerated from the seed by `cloud()`; only the reference's outputs are stored)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT]

CASES = {"uniform_20k": (20000, 1, "uniform"), "clustered_30k": (30000, 2, "clustered"), "tiny_5": (5, 3, "uniform"),
         "dups_3k": (3000, 4, "dups")}


def cloud(n, seed, kind):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return (rng.uniform(-3, 3, size=(n, 3))).astype(np.float32)
    if kind == "clustered":      # SfM-like: dense clusters of very different scale + sparse far outliers
        c = rng.normal(size=(40, 3)) * 5
        s = 10.0 ** rng.uniform(-3, 0, size=40)
        k = rng.integers(0, 40, size=n)
        p = c[k] + rng.normal(size=(n, 3)) * s[k][:, None]
        p[: n // 100] = rng.normal(size=(n // 100, 3)) * 300
        return p.astype(np.float32)
    p = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    p[n // 2:] = p[: n - n // 2]            # every point of the first half has an exact duplicate
    return p


def main():
    import torch
    from oracle import ref_knn
    assert ref_knn.available()
    for name, (n, seed, kind) in CASES.items():
        pts = cloud(n, seed, kind)
        out = ref_knn.distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
        np.savez_compressed(os.path.join(HERE, f"knn_{name}.npz"), mean_dist2=out, n=n, seed=seed, kind=kind)
        print(name, out[:3], float(out.max()))


if __name__ == "__main__":
    main()
