"""TEST INFRASTRUCTURE -- numpy restatement of ``GaussianModel.compute_3D_filter`` (wildgaussians/method.py:1140-1190):
the per-Gaussian 3D low-pass filter size of Mip-Splatting from the closest camera that sees the Gaussian.

Pinned by tests/test_filter3d.py against tests/golden/filter3d_*.npz, which tests/golden/make_golden_filter3d.py produced by
running the reference's own, unmodified method on the CPU.  Only tests / smoke / bench may import this file."""
from __future__ import annotations

import numpy as np

f32 = np.float32


def camera_matrices(poses):
    """method.py:1153-1158: camera-to-world 3x4 pose -> (R, T) of `xyz @ R + T`."""
    pose = np.concatenate([np.array(poses), np.array([[0, 0, 0, 1]], dtype=np.asarray(poses).dtype)], axis=0)
    w2c = np.linalg.inv(pose)
    return np.transpose(w2c[:3, :3]).astype(f32), w2c[:3, 3].astype(f32)


def compute_3d_filter(xyz, cameras):
    """xyz [P,3] fp32; cameras: objects with .poses (3x4), .intrinsics (fx, fy, cx, cy), .image_sizes (W, H).
    Returns filter_3D [P] fp32 (the reference registers it as [P,1])."""
    xyz = np.asarray(xyz, dtype=f32)
    P = xyz.shape[0]
    distance = np.full((P,), 100000.0, dtype=f32)                          # :1142
    seen = np.zeros((P,), dtype=bool)                                      # :1143
    focal = 0.0
    for cam in cameras:
        fx, fy = cam.intrinsics[0], cam.intrinsics[1]
        W, H = cam.image_sizes
        R, T = camera_matrices(cam.poses)
        # :1165  xyz @ R + T, one fp32 rounding per product / sum, products accumulated left to right
        c = [(((xyz[:, 0] * R[0, j]).astype(f32) + (xyz[:, 1] * R[1, j]).astype(f32)).astype(f32)
              + (xyz[:, 2] * R[2, j]).astype(f32)).astype(f32) + T[j] for j in range(3)]
        deep = c[2] > f32(0.2)                                             # :1168
        z = np.maximum(c[2], f32(0.001))                                   # :1172
        px = ((c[0] / z).astype(f32) * f32(fx)).astype(f32) + f32(W / 2.0)  # :1174
        py = ((c[1] / z).astype(f32) * f32(fy)).astype(f32) + f32(H / 2.0)  # :1175
        inside = (px >= f32(-0.15 * W)) & (px <= f32(W * 1.15)) & (py >= f32(-0.15 * H)) & (py <= f32(1.15 * H))   # :1177
        ok = deep & inside
        distance[ok] = np.minimum(distance[ok], z[ok])                     # :1182
        seen |= ok
        if focal < fx:
            focal = fx                                                     # :1184-1185
    distance[~seen] = distance[seen].max()                                 # :1187
    return ((distance / f32(focal)).astype(f32) * f32(0.2 ** 0.5)).astype(f32)   # :1189
