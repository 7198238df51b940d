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


def _io_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import parallel
        g = torch.Generator().manual_seed(5)
        ok = True
        for shape in ((1001, 3), (7,), (3, 50, 70), (4, 4)):
            host = torch.randn(*shape, generator=g)              # same data on every rank
            full = parallel.upload_sharded(host, torch.device("cpu"))
            ok = ok and torch.equal(full, host)
            out = torch.full(shape, float("nan"))
            parallel.download_sharded(full * 2, out)
            # every rank wrote exactly its chunk; together the chunks tile the tensor
            flat = out.reshape(-1)
            n = flat.numel(); chunk = (n + world - 1) // world
            lo, hi = min(n, rank * chunk), min(n, (rank + 1) * chunk)
            mine = torch.zeros(n, dtype=torch.bool); mine[lo:hi] = True
            ok = ok and bool(torch.equal(flat[mine], (host.reshape(-1) * 2)[mine])) and bool(torch.isnan(flat[~mine]).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sharded_host_io_world2():
    """parallel.upload_sharded / download_sharded (the multi-GPU e2e path: every rank moves 1/world of each tensor) on
    gloo, world_size 2: the all-gathered tensor equals the host tensor; the ranks' downloads tile the result."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_io_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
