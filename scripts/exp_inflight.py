"""Experiment: throughput with D frames in flight (D renderers, D streams, one shared point cloud)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench                      # noqa: E402
import websplat_b200 as ws        # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
ctx = ws.Context(0)
cloud, W, H, views = bench.make_workload(name)
gen = ws.GenericGaussianPointCloud(cloud["gaussians"], cloud["sh_coefs"], cloud["sh_deg"], cloud["num_points"],
                                   ws.Aabb(cloud["aabb_min"], cloud["aabb_max"]), cloud["center"], compressed=cloud["compressed"],
                                   covars=cloud.get("covars"), quantization=cloud.get("quantization"))
pc = ws.PointCloud.new(ctx, gen)
fargs = [bench.frame_args(ws, cloud, v, W, H) for v in views]
for D in (1, 2, 3, 4):
    rs = []
    for d in range(D):
        r = ws.GaussianRenderer.new(ctx, ws.FORMAT_RGBA16_FLOAT, cloud["sh_deg"], cloud["compressed"])
        r.set_pair_capacity(min(max(8 * cloud["num_points"], 1 << 22), (1 << 30) - 1))
        r.set_timing(False)
        rs.append((r, torch.cuda.Stream(), torch.empty((H, W, 4), dtype=torch.float16, device="cuda")))

    def frame(i):
        r, s, t = rs[i % D]
        r.prepare(s, pc, fargs[i % len(fargs)])
        r.render(t, pc, stream=s)

    for i in range(8 * D):
        frame(i)
    torch.cuda.synchronize()
    K = 360
    t0 = time.perf_counter()
    for i in range(K):
        frame(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("frames in flight %d: %.1f frames/s (%.3f ms/frame)" % (D, K / dt, dt / K * 1e3), flush=True)
    for r, _, _ in rs:
        r.close()
