#!/usr/bin/env python
"""Debug aid (GPU box, N GPUs): which outputs of the tile-row sharded path differ from the single-GPU result, per rank / mode.
usage: python tools/debug_multi.py WORLD MODE [P W H]"""
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, port, mode, P, W, H):
    os.environ["GSR_PEER_REDUCE"] = str(mode)
    for p in (ROOT, os.path.join(ROOT, "wild-gaussians_b200")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import parallel
    import synthetic
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    scene = synthetic.make_scene(P=P, W=W, H=H, sh_degree=None, seed=77)
    d = synthetic.to_device(scene, dev)
    st = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=d["tanfovx"], tanfovy=d["tanfovy"], kernel_size=d["kernel_size"],
        subpixel_offset=d["subpixel_offset"], bg=d["bg"], scale_modifier=1.0, viewmatrix=d["viewmatrix"],
        projmatrix=d["projmatrix"], sh_degree=0, campos=d["campos"], prefiltered=False, debug=False, return_accumulation=True)
    for rep in range(3):
        res = {}
        for name, cls in (("single", GaussianRasterizer), ("sharded", parallel.ShardedGaussianRasterizer)):
            leaves = {k: d[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "colors_precomp")}
            m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
            img, radii, acc = cls(st)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                     colors_precomp=leaves["colors_precomp"], scales=leaves["scales"], rotations=leaves["rotations"])
            (img * d["dL_dpix"]).sum().backward()
            res[name] = dict(img=img.detach(), radii=radii, acc=acc, m2d=m2d.grad, **{k: v.grad for k, v in leaves.items()})
        torch.cuda.synchronize()
        a, b = res["single"], res["sharded"]
        bad_rows = (a["img"] != b["img"]).any(dim=0).any(dim=1).nonzero().flatten().tolist()
        bad_acc = (a["acc"] != b["acc"]).any(dim=1).nonzero().flatten().tolist()
        worst = max(float((a[k] - b[k]).abs().max()) / (float(a[k].abs().max()) + 1e-30) for k in ("means3D", "opacities", "scales", "rotations", "colors_precomp", "m2d"))
        print(f"mode {mode} world {world} rank {rank} rep {rep}: img rows differing {len(bad_rows)} {bad_rows[:4]}..{bad_rows[-2:]}, "
              f"acc rows {len(bad_acc)} {bad_acc[:3]}, radii diff {int((a['radii'] != b['radii']).sum())}, grad rel {worst:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    world, mode = int(sys.argv[1]), int(sys.argv[2])
    P, W, H = (int(x) for x in sys.argv[3:6]) if len(sys.argv) > 5 else (120_000, 640, 368)
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(worker, args=(world, port, mode, P, W, H), nprocs=world, join=True)
