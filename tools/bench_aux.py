#!/usr/bin/env python
"""Timings of the SURVEY 8f-4 rows that are not part of a train step (GPU box; one JSON line):
  * distCUDA2 (simple-knn): this repo's csrc/knn.cu vs the compiled unmodified reference (oracle/_ref/libsimpleknn_ref.so),
    bit-exactness asserted, on uniform and clustered clouds;
  * compute_3D_filter: csrc/filter3d.cu vs the unmodified method.py statements on the same device.
usage: python tools/bench_aux.py [P_knn] [P_filter] [C_cameras]"""
import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wild-gaussians_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import make_golden_knn as mg  # noqa: E402
import make_golden_filter3d as mf  # noqa: E402


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(t))


def main():
    P_knn = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    P_f = int(sys.argv[2]) if len(sys.argv) > 2 else 3_000_000
    C = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    from simple_knn._C import distCUDA2
    from oracle import ref_knn
    out = {"knn": {}, "filter3d": {}}
    for kind in ("uniform", "clustered"):
        pts = torch.from_numpy(mg.cloud(P_knn, 41, kind)).cuda()
        r = {"P": P_knn, "ours_ms": timed(lambda: distCUDA2(pts))}
        if ref_knn.available():
            r["reference_ms"] = timed(lambda: ref_knn.distCUDA2(pts), n=2)
            a, b = distCUDA2(pts), ref_knn.distCUDA2(pts)
            r["bit_exact"] = bool(torch.equal(a, b))
            r["mismatches"] = int((a != b).sum())
        out["knn"][kind] = r
    import wg_harness as wh
    import wildgaussians_fused as wf
    m, Config = wh.import_method()
    if m is not None:
        cfg = Config(source_path="", model_path="", uncertainty_mode="disabled")
        model = m.GaussianModel(cfg, training_setup=False).cuda()
        model._resize_parameters(P_f)
        with torch.no_grad():
            model.xyz.copy_(torch.randn(P_f, 3, generator=torch.Generator().manual_seed(3)) * 2.5)
        cams = mf.make_cameras(C, 5)
        r = {"P": P_f, "cameras": C, "method_py_ms": timed(lambda: type(model).compute_3D_filter(model, cams), n=2)}
        want = model.filter_3D.clone()
        r["ours_ms"] = timed(lambda: wf.compute_3D_filter(model, cams))
        rel = ((model.filter_3D - want).abs() / want.abs())
        r["entries_beyond_2e-6"] = int((rel > 2e-6).sum())
        out["filter3d"] = r
    print(json.dumps(out))


if __name__ == "__main__":
    main()
