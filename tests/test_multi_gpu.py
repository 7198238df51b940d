"""GPU test of the tile-row sharded path over NCCL (needs >= 2 GPUs; skipped otherwise): every rank must
obtain the single-GPU image bit for bit and the single-GPU gradients within the atomics tolerance."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, peer_mode=0):
    import sys
    os.environ["GSR_PEER_REDUCE"] = str(peer_mode)     # read when parallel.py is imported
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "wild-gaussians_b200"), os.path.join(root, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import parallel
        import synthetic
        from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
        scene = synthetic.make_scene(P=120_000, W=640, H=368, sh_degree=None, seed=77)
        d = synthetic.to_device(scene, dev)
        st = GaussianRasterizationSettings(
            image_height=d["image_height"], image_width=d["image_width"], tanfovx=d["tanfovx"], tanfovy=d["tanfovy"],
            kernel_size=d["kernel_size"], subpixel_offset=d["subpixel_offset"], bg=d["bg"], scale_modifier=1.0,
            viewmatrix=d["viewmatrix"], projmatrix=d["projmatrix"], sh_degree=0, campos=d["campos"],
            prefiltered=False, debug=False, return_accumulation=True)
        res = {}
        for name, cls in (("single", GaussianRasterizer), ("sharded", parallel.ShardedGaussianRasterizer)):
            leaves = {k: d[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "colors_precomp")}
            m2d = torch.zeros_like(leaves["means3D"], requires_grad=True)
            img, radii, acc = cls(st)(means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"],
                                     colors_precomp=leaves["colors_precomp"], scales=leaves["scales"],
                                     rotations=leaves["rotations"])
            (img * d["dL_dpix"]).sum().backward()
            res[name] = dict(img=img.detach(), radii=radii, acc=acc, m2d=m2d.grad, **{k: v.grad for k, v in leaves.items()})
        torch.cuda.synchronize()
        ok = torch.equal(res["single"]["img"], res["sharded"]["img"]) and torch.equal(res["single"]["radii"], res["sharded"]["radii"]) \
            and torch.equal(res["single"]["acc"], res["sharded"]["acc"])
        worst = 0.0
        for k in ("means3D", "opacities", "scales", "rotations", "colors_precomp", "m2d"):
            a, b = res["single"][k], res["sharded"][k]
            worst = max(worst, float((a - b).abs().max()) / (float(a.abs().max()) + 1e-30))
        q.put((rank, ok, worst))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("peer_mode", [0, 1, 2, 3], ids=["nccl_allreduce", "peer_red", "multicast_red", "pull"])
def test_sharded_equals_single_gpu(peer_mode):
    """peer_mode 0: NCCL all-reduce of the partial gradients; 1 / 2: the reduction fused into the backward composite
    through peer pointers / the NVSwitch multicast address (symmetric memory); 3: local sums + marks, every rank pulls the
    marked rows of the others (no remote atomics)."""
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    # every visible GPU takes part (8 on the scaling box): the 8-band partition has 2-3 tile rows per rank here
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, peer_mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok, worst in res:
        assert ok, f"rank {rank}: sharded image / radii / accumulation differ from the single-GPU result"
        assert worst < 1e-3, f"rank {rank}: sharded gradients differ ({worst})"
