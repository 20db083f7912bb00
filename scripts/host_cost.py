"""Host-side (CPU) cost of submitting one frame: how many microseconds the calling thread spends per frame in
prepare()+render() (CUDA-graph replay) and in the sharded single-call frame (world 1: same code path, ~25 launches).
A tiny cloud keeps the GPU work negligible, so the loop is submission-bound.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                      # noqa: E402
import websplat_b200 as ws        # noqa: E402
from helpers import make_args, make_generic   # noqa: E402


def main():
    ctx = ws.Context(0)
    n, W, H = 20000, 640, 360
    cloud = ws.synth.make_cloud(n, 3)
    pc = ws.PointCloud.new(ctx, make_generic(ws, cloud))
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    args = [make_args(ws, cloud, *ws.synth.orbit_camera(az), W, H, fovx, fovy) for az in (0.0, 90.0, 180.0, 270.0)]
    out = {}
    r = ws.GaussianRenderer.new(ctx, ws.FORMAT_RGBA16_FLOAT, 3, False)
    r.set_timing(False)
    t = torch.empty((H, W, 4), dtype=torch.float16, device="cuda")
    st = torch.cuda.Stream()
    for graphs in (True, False):
        r.set_cuda_graphs(graphs)
        for i in range(50):
            r.prepare(st, pc, args[i % 4]); r.render(t, pc, stream=st)
        torch.cuda.synchronize()
        K = 2000
        t0 = time.perf_counter()
        for i in range(K):
            r.prepare(st, pc, args[i % 4]); r.render(t, pc, stream=st)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out["plain_graphs_%s" % ("on" if graphs else "off")] = {"submit_us_per_frame": 1e6 * (t1 - t0) / K, "total_us_per_frame": 1e6 * (t2 - t0) / K}
    for depth in (1, 2):
        pipe = ws.ShardedPipeline(ws, ctx, ws.FORMAT_RGBA16_FLOAT, 3, False, pc, n, (W, H), depth=depth)
        for i in range(50):
            pipe.frame_peer(args[i % 4])
        pipe.synchronize(); torch.cuda.synchronize()
        K = 2000
        t0 = time.perf_counter()
        for i in range(K):
            pipe.frame_peer(args[i % 4])
        t1 = time.perf_counter()
        pipe.synchronize(); torch.cuda.synchronize()
        t2 = time.perf_counter()
        out["sharded_world1_depth%d" % depth] = {"submit_us_per_frame": 1e6 * (t1 - t0) / K, "total_us_per_frame": 1e6 * (t2 - t0) / K}
        del pipe
    print(json.dumps(out))


if __name__ == "__main__":
    main()
