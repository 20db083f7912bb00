"""torchrun worker for compute-sanitizer runs of the sharded frame (synccheck / memcheck / racecheck): a small cloud, a few
single-call frames (mailbox flags in peer memory, CUDA-graph replay, occlusion split on the last one) on every rank,
result compared with the 1-GPU frame on rank 0.  Launched as
    python -m torch.distributed.run --nproc-per-node 2 --no-python compute-sanitizer --tool synccheck python tests/sharded_sync_worker.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch                            # noqa: E402
import torch.distributed as dist        # noqa: E402
import websplat_b200 as ws              # noqa: E402
from helpers import make_args, make_generic   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ws.Context(local)
    n, W, H = 20000, 640, 360
    cloud = ws.synth.make_cloud(n, 77)
    pc = ws.PointCloud.new(ctx, make_generic(ws, ws.shard_cloud(cloud, rank, world)))
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    args = make_args(ws, cloud, *ws.synth.orbit_camera(33.0), W, H, fovx, fovy)
    pipe = ws.ShardedPipeline(ws, ctx, ws.FORMAT_RGBA16_FLOAT, 3, False, pc, n, (W, H), depth=2)
    host = torch.zeros((H, W, 4), dtype=torch.float16).pin_memory() if rank == 0 else None
    for i in range(6):
        for s in pipe.slots:
            s.r.set_occlusion_split(i >= 4)
        pipe.frame_peer(args, host=host)
    pipe.synchronize()
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        full = ws.PointCloud.new(ctx, make_generic(ws, cloud))
        plain = ws.GaussianRenderer.new(ctx, ws.FORMAT_RGBA16_FLOAT, 3, False)
        plain.prepare(None, full, args)
        ref = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")
        plain.render(ref, full)
        torch.cuda.synchronize()
        ok = torch.equal(host, ref.cpu())
        print("sanitized sharded frame identical=%s" % ok, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
