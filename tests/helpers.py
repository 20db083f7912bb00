"""Shared helpers for the parity tests."""
import numpy as np


def make_generic(ws, cloud):
    return ws.GenericGaussianPointCloud(
        cloud["gaussians"], cloud["sh_coefs"], cloud["sh_deg"], cloud["num_points"],
        ws.Aabb(cloud["aabb_min"], cloud["aabb_max"]), cloud["center"], compressed=cloud["compressed"],
        covars=cloud.get("covars"), quantization=cloud.get("quantization"))


def make_args(ws, cloud, pos, rot, W, H, fovx, fovy, **kw):
    cam = ws.PerspectiveCamera(pos, rot, ws.PerspectiveProjection(fovx, fovy, 0.1, 100.0))
    cam.fit_near_far(ws.Aabb(cloud["aabb_min"], cloud["aabb_max"]))     # as every reference caller does
    return ws.SplattingArgs(cam, (W, H), **kw)


def f16_ordered(bits):
    """map f16 bit patterns to integers that are monotone in the value (for ulp distances)."""
    i = bits.astype(np.int32)
    return np.where(i & 0x8000, 0x8000 - (i & 0x7fff), (i & 0x7fff) + 0x8000)


def oracle_pairs(orc, splats, order, W, H):
    """(tile, slot) pair list sorted by (tile, draw order) from the oracle's stage outputs."""
    rects, P = orc.tile_rects(splats, W, H)
    tx = (W + 15) // 16
    tiles, slots = [], []
    for slot in order:                          # draw order = ascending key, ties by slot
        x0, y0, x1, y1 = rects[slot]
        if x1 < x0:
            continue
        ys, xs = np.mgrid[y0:y1 + 1, x0:x1 + 1]
        t = (ys * tx + xs).reshape(-1)
        tiles.append(t); slots.append(np.full(t.size, slot, np.uint32))
    if not tiles:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32), rects, 0
    tiles = np.concatenate(tiles).astype(np.uint32); slots = np.concatenate(slots)
    o = np.argsort(tiles, kind="stable")
    return tiles[o], slots[o], rects, P


def image_close(img, ref, sens, atol=2e-3):
    """|img - ref| <= atol + sens per pixel (sens = oracle's bound on discard-threshold flips)."""
    d = np.abs(img.astype(np.float64) - ref.astype(np.float64)).max(axis=2)
    lim = atol + (sens if sens is not None else 0.0)
    return d, (d <= lim)
