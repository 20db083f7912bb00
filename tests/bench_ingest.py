"""(test infrastructure: it times the CPU oracle next to the product, so it lives under tests/)
Times the section 8(f) ingest rows: .ply / .npz file image in host memory -> resident GPU layouts,
next to the CPU oracle's conversion of the same arrays.  Prints one JSON line per format."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import websplat_b200 as ws           # noqa: E402
from oracle import oracle as orc     # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
    ctx = ws.Context(0)
    v = ws.synth.ply_vertices(n, 1, 3)
    img = ws.synth.ply_bytes(v, 3)
    ws.PointCloud.from_ply(ctx, img).close()                     # warm-up: context, allocator
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); pc = ws.PointCloud.from_ply(ctx, img); ts.append(time.perf_counter() - t0); pc.close()
    t0 = time.perf_counter(); orc.ply_convert(v, 3); t_cpu = time.perf_counter() - t0
    print(json.dumps(dict(format="ply", points=n, file_mb=len(img) / 1e6, gpu_ingest_ms=min(ts) * 1e3,
                          gpoints_per_s=n / min(ts) / 1e9, h2d_gbps=len(img) / min(ts) / 1e9,
                          oracle_ms=t_cpu * 1e3, oracle_threads=orc.num_threads())))
    del v, img
    a = ws.synth.c3dgs_arrays(n, 1, 3, codebook=4096)
    ws.PointCloud.from_npz(ctx, a).close()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); pc = ws.PointCloud.from_npz(ctx, a); ts.append(time.perf_counter() - t0); pc.close()
    t0 = time.perf_counter(); orc.c3dgs_convert(a); t_cpu = time.perf_counter() - t0
    print(json.dumps(dict(format="npz(arrays)", points=n, gpu_ingest_ms=min(ts) * 1e3, gpoints_per_s=n / min(ts) / 1e9,
                          oracle_ms=t_cpu * 1e3)))


if __name__ == "__main__":
    main()
