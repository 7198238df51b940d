"""CPU tests of the multi-GPU host logic with the gloo backend, world_size 2: band partition,
image all-gather, all-reduce of the per-Gaussian partials.  Each rank produces its band with the CPU oracle
(tests may use it); the collectives under test are the product's (wild-gaussians_b200/parallel.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parallel
import synthetic


def test_partition_tile_rows():
    assert parallel.partition_tile_rows(68, 8) == [(0, 9), (9, 17), (17, 26), (26, 34), (34, 43), (43, 51), (51, 60), (60, 68)] \
        or sum(b - a for a, b in parallel.partition_tile_rows(68, 8)) == 68
    for n, w in ((68, 8), (135, 8), (7, 2), (3, 4), (0, 2), (1, 1)):
        bands = parallel.partition_tile_rows(n, w)
        assert len(bands) == w and bands[0][0] == 0 and bands[-1][1] == n
        assert all(bands[i][1] == bands[i + 1][0] for i in range(w - 1))
        assert all(b >= a for a, b in bands)
        if n >= w:
            sizes = [b - a for a, b in bands]
            assert max(sizes) - min(sizes) <= 1
    # weighted: the heavy rows are split off
    bands = parallel.partition_tile_rows(6, 2, weights=[10, 1, 1, 1, 1, 1])
    assert bands == [(0, 1), (1, 6)]
    assert parallel.band_pixel_rows((2, 5), 70) == (32, 70)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cpu_oracle
        cpu_oracle.set_num_threads(2)
        scene = synthetic.make_scene(P=1200, W=90, H=75, sh_degree=None, seed=5, scale_range=(0.01, 0.1))
        H, W = 75, 90
        rows = (H + 15) // 16
        bands = parallel.partition_tile_rows(rows, world)
        st = cpu_oracle.forward(scene, tile_rows=bands[rank])
        img = torch.from_numpy(np.concatenate([st["out_color"], st["final_T"][None]], axis=0))
        r0, r1 = parallel.band_pixel_rows(bands[rank], H)
        mask = torch.zeros_like(img); mask[:, r0:r1] = 1
        img = img * mask + (1 - mask) * 123.0          # rows outside the band hold garbage
        full = parallel.gather_image_bands(img, bands)
        g = cpu_oracle.backward(st, scene["dL_dpix"])
        acc = torch.from_numpy(g["acc"].astype(np.float32))
        parallel.reduce_partials(acc)
        ref = cpu_oracle.forward(scene)
        gref = cpu_oracle.backward(ref, scene["dL_dpix"])
        ok_img = np.array_equal(full[:3].numpy(), ref["out_color"]) and np.array_equal(full[3].numpy(), ref["final_T"])
        err = float(np.abs(acc.numpy() - gref["acc"]).max() / max(1.0, np.abs(gref["acc"]).max()))
        q.put((rank, ok_img, err, int(st["num_rendered"]), int(ref["num_rendered"])))
    finally:
        dist.destroy_process_group()


def test_gather_and_reduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sum(r[3] for r in res) == res[0][4], "bands do not partition the instance list"
    for rank, ok_img, err, _, _ in res:
        assert ok_img, f"rank {rank}: gathered image differs from the single-rank image"
        assert err < 1e-6, f"rank {rank}: reduced partials differ ({err})"
