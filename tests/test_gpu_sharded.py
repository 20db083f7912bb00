"""GPU tests of the sharded path.  world == 1 exercises the whole begin/exchange/finish/band sequence
on one GPU (the exchange then stores into local memory); the 2-GPU test runs under
`gpurun --gpus 2` and checks that the 2-GPU frame is BIT-IDENTICAL to the 1-GPU frame."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import make_args, make_generic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_world1_equals_plain(ws, orc, ctx):
    import torch
    cloud = ws.synth.make_cloud(50000, 31)
    W, H = 640, 360
    pos, rot = ws.synth.orbit_camera(110.0)
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    args = make_args(ws, cloud, pos, rot, W, H, fovx, fovy)
    pc = ws.PointCloud.new(ctx, make_generic(ws, cloud))
    plain = ws.GaussianRenderer.new(ctx, ws.FORMAT_RGBA32_FLOAT, 3, False)
    plain.set_occlusion_split(False)                     # num_pairs is compared below; sharded frames never split
    plain.prepare(None, pc, args)
    t = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    plain.render(t, pc)
    torch.cuda.synchronize()
    sh = ws.ShardedRenderer(ws, ctx, ws.FORMAT_RGBA32_FLOAT, 3, False, pc, cloud["num_points"], (W, H))
    for _ in range(2):                                   # twice: buffers are reused frame to frame
        img = sh.frame(args)
        torch.cuda.synchronize()
        assert torch.equal(img, t)
    host = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
    sh.frame_to_root(args, host=host)
    torch.cuda.synchronize()
    assert torch.equal(host, t.cpu())
    for _ in range(3):                                   # host-collective-free variant, epochs advance
        host.zero_()
        sh.frame_peer(args, host=host)
        torch.cuda.synchronize()
        assert torch.equal(host, t.cpu())
    st = sh.stats()
    # only splats that touch at least one tile are routed, so the received count can be below V
    assert st["num_visible"] <= plain.stats()["num_visible"] and st["num_pairs"] == plain.stats()["num_pairs"]
    # two frames in flight through the gated path (world 1: the gates wait on this rank's own flags)
    pipe = ws.ShardedPipeline(ws, ctx, ws.FORMAT_RGBA32_FLOAT, 3, False, pc, cloud["num_points"], (W, H), depth=2)
    hosts = [torch.zeros((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(4)]
    for h in hosts:
        pipe.frame_peer(args, host=h)
    pipe.synchronize()
    torch.cuda.synchronize()
    for h in hosts:
        assert torch.equal(h, t.cpu())
    assert pipe.rebalance() == [0, (H + 15) // 16]


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpus_bit_identical_to_one():
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "sharded_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "SHARDED_OK" in p.stdout
