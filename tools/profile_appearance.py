#!/usr/bin/env python
"""One forward + backward of the fused colour op (csrc/appearance.cu) on P rows, for ncu:
  ncu --set full --clock-control none --import-source on -k regex:appearance_ -o gpurun_out/r2_appearance \
      python tools/profile_appearance.py [P]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "wild-gaussians_b200")]
import fused_colors as fc  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(11)
R = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(dev)
leaves = [t.requires_grad_(True) for t in (R(P, 3), R(P, 45, scale=0.3), R(P, 24, scale=0.5), R(32, scale=0.5), R(P, 3, scale=2.0))]
torch.manual_seed(3)
mlp = torch.nn.Sequential(torch.nn.Linear(59, 128), torch.nn.ReLU(), torch.nn.Linear(128, 128), torch.nn.ReLU(),
                          torch.nn.Linear(128, 6)).to(dev)
with torch.no_grad():
    mlp[4].bias[3:] = 80.0
campos = torch.tensor([0.1, -0.2, 0.3], device=dev)
for _ in range(2):
    raw, toned = fc.fused_colors(*leaves[:4], mlp, leaves[4], campos, 3)
    torch.autograd.backward([raw, toned], [R(P, 3), R(P, 3)])
torch.cuda.synchronize()
print("ok", float(toned.mean()))
