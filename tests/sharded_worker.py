"""torchrun worker for tests/test_gpu_sharded.py: renders the same frames on G GPUs (sharded) and on
one GPU (rank 0, plain path) and requires bit-identical images."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np                      # noqa: E402
import torch                            # noqa: E402
import torch.distributed as dist        # noqa: E402
import websplat_b200 as ws              # noqa: E402
from helpers import make_args, make_generic   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ws.Context(local)
    ok = True
    for n, W, H, fmt in ((60000, 800, 600, ws.FORMAT_RGBA32_FLOAT), (200000, 1200, 799, ws.FORMAT_RGBA16_FLOAT)):
        cloud = ws.synth.make_cloud(n, 1234 + n)
        shard = ws.shard_cloud(cloud, rank, world)
        pc = ws.PointCloud.new(ctx, make_generic(ws, shard))
        sh = ws.ShardedRenderer(ws, ctx, fmt, 3, False, pc, n, (W, H))
        fovx, fovy = ws.synth.fov_for_viewport(W, H)
        for az in (0.0, 95.0, 250.0):
            pos, rot = ws.synth.orbit_camera(az)
            args = make_args(ws, cloud, pos, rot, W, H, fovx, fovy)
            img = sh.frame(args, clear=(0.1, 0.2, 0.3, 0.5))
            torch.cuda.synchronize()
            if rank == 0:
                full = ws.PointCloud.new(ctx, make_generic(ws, cloud))
                plain = ws.GaussianRenderer.new(ctx, fmt, 3, False)
                plain.prepare(None, full, args)
                ref = torch.empty_like(img)
                plain.render(ref, full, (0.1, 0.2, 0.3, 0.5))
                torch.cuda.synchronize()
                same = torch.equal(ref, img)
                ok = ok and same
                print("n=%d %dx%d az=%.0f identical=%s V=%d" % (n, W, H, az, same, sh.stats()["num_visible"]), flush=True)
            # the variant whose band gather is fused into the compositor (peer stores into rank 0's frame)
            host = torch.zeros_like(img, device="cpu").pin_memory() if rank == 0 else None
            sh.frame_to_root(args, clear=(0.1, 0.2, 0.3, 0.5), root=0, host=host)
            torch.cuda.synchronize()
            if rank == 0:
                same2 = torch.equal(host, ref.cpu())
                ok = ok and same2
                print("   to_root identical=%s" % same2, flush=True)
            # and the variant without any host-side collective (mailbox flags in peer memory), several frames back to back;
            # the last one with the occlusion split forced on inside every band (automatic only from 2 M points per rank)
            for rep in range(3):
                if rank == 0:
                    host.zero_()
                sh.r.set_occlusion_split(rep == 2)
                sh.frame_peer(args, clear=(0.1, 0.2, 0.3, 0.5), root=0, host=host)
            sh.r.set_occlusion_split(None)
            torch.cuda.synchronize()
            if rank == 0:
                same3 = torch.equal(host, ref.cpu())
                ok = ok and same3
                print("   peer-mailbox identical=%s" % same3, flush=True)
            dist.barrier()
        # two frames in flight (gate kernels) over cost-balanced bands: still bit-identical, for several views in a row
        pipe = ws.ShardedPipeline(ws, ctx, fmt, 3, False, pc, n, (W, H), depth=2)
        views = [ws.synth.orbit_camera(az) for az in (10.0, 130.0, 200.0, 340.0)]
        vargs = [make_args(ws, cloud, p_, r_, W, H, fovx, fovy) for p_, r_ in views]
        for round_ in range(2):
            hosts = [torch.zeros((H, W, 4), dtype=img.dtype).pin_memory() for _ in vargs] if rank == 0 else [None] * len(vargs)
            for a_, h_ in zip(vargs, hosts):
                pipe.frame_peer(a_, clear=(0.1, 0.2, 0.3, 0.5), root=0, host=h_)
            pipe.synchronize()
            torch.cuda.synchronize()
            if rank == 0:
                full = ws.PointCloud.new(ctx, make_generic(ws, cloud))
                plain = ws.GaussianRenderer.new(ctx, fmt, 3, False)
                for a_, h_ in zip(vargs, hosts):
                    plain.prepare(None, full, a_)
                    ref = torch.empty((H, W, 4), dtype=img.dtype, device="cuda")
                    plain.render(ref, full, (0.1, 0.2, 0.3, 0.5))
                    torch.cuda.synchronize()
                    same4 = torch.equal(h_, ref.cpu())
                    ok = ok and same4
                print("   pipeline depth 2, bands %s identical=%s" % (pipe.bands, ok), flush=True)
            dist.barrier()
            new = pipe.rebalance()                       # second round runs on the re-cut bands
            assert new == pipe.slots[1].bands
        dist.barrier()
    # cfg5-style cloud (BASELINE.json configs[4]): every rank GENERATES only its own shard (seeded by rank, splat size of the
    # whole cloud); the union is the cloud.  Rank 0 rebuilds the union, renders it on one GPU and compares bit for bit.
    n_all, W, H = 800000, 1920, 1080
    lo_i, hi_i = (n_all * rank) // world, (n_all * (rank + 1)) // world
    mine = ws.synth.make_cloud(hi_i - lo_i, 4242 + 7919 * rank, density_n=n_all)
    lo = torch.tensor(mine["aabb_min"], device="cuda"); hi = torch.tensor(mine["aabb_max"], device="cuda")
    csum = torch.tensor(mine["center"].astype(np.float64) * (hi_i - lo_i), device="cuda")
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(csum)
    meta = dict(aabb_min=lo.cpu().numpy(), aabb_max=hi.cpu().numpy(), center=(csum / n_all).cpu().numpy().astype(np.float32))
    mine = dict(mine, **meta)
    whole_meta = dict(mine, num_points=n_all)
    pc = ws.PointCloud.new(ctx, make_generic(ws, mine))
    pipe = ws.ShardedPipeline(ws, ctx, ws.FORMAT_RGBA16_FLOAT, 3, False, pc, n_all, (W, H), depth=2)
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    vargs = [make_args(ws, whole_meta, *ws.synth.orbit_camera(az), W, H, fovx, fovy) for az in (0.0, 120.0, 240.0)]
    hosts = [torch.zeros((H, W, 4), dtype=torch.float16).pin_memory() for _ in vargs] if rank == 0 else [None] * len(vargs)
    for a_, h_ in zip(vargs, hosts):
        pipe.frame_peer(a_, root=0, host=h_)
    pipe.synchronize()
    torch.cuda.synchronize()
    if rank == 0:
        parts = [mine if q == 0 else ws.synth.make_cloud((n_all * (q + 1)) // world - (n_all * q) // world, 4242 + 7919 * q, density_n=n_all)
                 for q in range(world)]
        whole = dict(whole_meta, gaussians=np.concatenate([p_["gaussians"] for p_ in parts]),
                     sh_coefs=np.concatenate([p_["sh_coefs"] for p_ in parts]))
        full = ws.PointCloud.new(ctx, make_generic(ws, whole))
        plain = ws.GaussianRenderer.new(ctx, ws.FORMAT_RGBA16_FLOAT, 3, False)
        for a_, h_ in zip(vargs, hosts):
            plain.prepare(None, full, a_)
            ref = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")
            plain.render(ref, full)
            torch.cuda.synchronize()
            same5 = torch.equal(h_, ref.cpu())
            ok = ok and same5
            print("cfg5-style: %d ranks x own shard (n=%d) identical=%s" % (world, n_all, same5), flush=True)
    dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and int(flag.item()) == 1:
        print("SHARDED_OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
