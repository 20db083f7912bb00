"""CPU tests (gloo, world_size 2) of the host-side logic of the sharded path: band partition, cloud
sharding, the G x G count matrix exchange and the receive offsets the exchange kernel uses."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_row_bands_cover_the_frame(ws):
    for H in (17, 600, 799, 1080, 2160):
        for G in (1, 2, 3, 4, 8):
            b = ws.tile_row_bands(H, G)
            ty = (H + 15) // 16
            assert b[0] == 0 and b[-1] == ty and all(b[i] <= b[i + 1] for i in range(G))
            assert max(b[i + 1] - b[i] for i in range(G)) - min(b[i + 1] - b[i] for i in range(G)) <= 1


def test_shard_cloud_partitions_gaussians(ws):
    cloud = ws.synth.make_cloud(1003, 5)
    parts = [ws.shard_cloud(cloud, r, 4) for r in range(4)]
    assert sum(p["num_points"] for p in parts) == 1003
    assert np.array_equal(np.concatenate([p["gaussians"] for p in parts]), cloud["gaussians"])
    assert np.array_equal(np.concatenate([p["sh_coefs"] for p in parts]), cloud["sh_coefs"])
    for p in parts:                                    # global metadata is kept
        assert np.array_equal(p["aabb_min"], cloud["aabb_min"]) and np.array_equal(p["center"], cloud["center"])


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import websplat_b200 as ws
    from oracle import oracle as orc
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    # every rank derives, from its shard of the oracle's stage-1 output, its row of the count matrix
    W, H = 320, 208
    cloud = ws.synth.make_cloud(4000, 77)
    pos, rot = ws.synth.orbit_camera(20.0)
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    zn, zf = orc.fit_near_far(pos, cloud["aabb_min"], cloud["aabb_max"])
    cam = orc.camera_uniform(pos, rot, fovx, fovy, zn, zf, W, H)
    shard = ws.shard_cloud(cloud, rank, world)
    splats, keys, src = orc.preprocess(shard, cam, orc.render_settings(cloud))
    rects, _ = orc.tile_rects(splats, W, H)
    bands = ws.tile_row_bands(H, world)
    row = torch.tensor([int(((rects[:, 3] >= rects[:, 1]) & (rects[:, 1] < bands[d + 1]) & (rects[:, 3] + 1 > bands[d])).sum())
                        for d in range(world)], dtype=torch.int32)
    rows = [torch.zeros(world, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(rows, row)
    matrix = torch.stack(rows).numpy()
    from websplat_b200.distributed import exchange_offsets
    off = exchange_offsets(matrix, rank)
    np.save(out % rank, np.concatenate([matrix.reshape(-1), off, [len(keys)]]))
    dist.destroy_process_group()


def test_count_matrix_and_offsets_gloo(ws, orc, tmp_path):
    import torch.multiprocessing as mp
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "r%d.npy")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    res = [np.load(out % r) for r in range(world)]
    m0, m1 = res[0][:4].reshape(2, 2), res[1][:4].reshape(2, 2)
    assert np.array_equal(m0, m1)                                  # every rank sees the same matrix
    assert np.array_equal(res[0][4:6], [0, 0])                     # rank 0 writes at the start of every destination
    assert np.array_equal(res[1][4:6], m0[0])                      # rank 1 after rank 0's records
    # every visible splat goes to at least one band; a splat straddling the boundary goes to both
    v = res[0][6] + res[1][6]
    assert m0.sum() >= v and m0.sum() <= 2 * v
    # the single-process oracle over the whole cloud gives the column sums
    W, H = 320, 208
    cloud = ws.synth.make_cloud(4000, 77)
    pos, rot = ws.synth.orbit_camera(20.0)
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    fr = orc.render_frame(cloud, pos, rot, W, H, fovx, fovy)
    rects, _ = orc.tile_rects(fr["splats"], W, H)
    bands = ws.tile_row_bands(H, world)
    for d in range(world):
        want = int(((rects[:, 3] >= rects[:, 1]) & (rects[:, 1] < bands[d + 1]) & (rects[:, 3] + 1 > bands[d])).sum())
        assert m0[:, d].sum() == want


def test_balanced_bands_properties(ws):
    """cost-balanced tile-row bands (distributed.balanced_bands): partition, monotone, at least one row each, better balance."""
    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 4, 8):
        for H in (600, 1080, 2160):
            ty = (H + 15) // 16
            rows = rng.uniform(0.2, 1.0, ty) * np.exp(-((np.arange(ty) - 0.4 * ty) / (0.3 * ty)) ** 2)     # a bump, like a real scene
            b = ws.tile_row_bands(H, world)
            imbalance = []
            for _ in range(4):
                loads = [rows[b[d]:b[d + 1]].sum() for d in range(world)]
                imbalance.append(max(loads) / (sum(loads) / world))
                nb = ws.balanced_bands(b, loads)
                assert nb[0] == 0 and nb[-1] == ty and len(nb) == world + 1
                assert all(nb[d + 1] > nb[d] for d in range(world))
                b = nb
            assert imbalance[-1] <= imbalance[0] + 1e-9
            if world in (2, 4, 8):
                assert imbalance[-1] < 1.0 + 1.5 * world / ty + 0.08          # within row granularity of perfect
    # degenerate inputs leave the bands alone / stay valid
    assert ws.balanced_bands([0, 5, 10], [0, 0]) == [0, 5, 10]
    assert ws.balanced_bands([0, 1, 2], [9, 1]) == [0, 1, 2]
    assert ws.balanced_bands([0, 2, 10], [100, 1]) == [0, 1, 10]
    assert ws.balanced_bands([0, 8, 10], [1, 100]) == [0, 9, 10]
