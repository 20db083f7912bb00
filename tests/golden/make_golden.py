"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).
The reference ships no fixtures of its own; these vectors freeze THIS repo's
restatement so the CUDA path and later oracle edits are compared against a fixed artefact."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import websplat_b200 as ws          # noqa: E402
from oracle import oracle as orc    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    n, seed, az, W, H = 3000, 42, 40.0, 160, 96
    cloud = ws.synth.make_cloud(n, seed)
    pos, rot = ws.synth.orbit_camera(az)
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    fr = orc.render_frame(cloud, pos, rot, W, H, fovx, fovy)
    _, P = orc.tile_rects(fr["splats"], W, H)
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), n=n, seed=seed, az=az, W=W, H=H,
                        splats=fr["splats"], keys=fr["keys"], order=fr["order"],
                        image_sub=fr["image"][::4, ::4], image=fr["image"].astype(np.float32), pairs=P)
    # the reference's sort KAT (gpu_rs.rs:295-331) as data
    keys = np.arange(8191, -1, -1, dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "sort_kat.npz"), keys_in=keys, keys_out=np.arange(8192, dtype=np.float32))


if __name__ == "__main__":
    main()
