"""compressed .npz ingest.  CPU: the oracle's restatement of NpzReader::read's array
post-processing against an independent numpy statement.  GPU: the ingest kernels against the oracle."""
import numpy as np
import pytest

from helpers import f16_ordered, make_args, make_generic


def _np_covars(a):
    """float64 statement of io/npz.rs:99-131,206-211."""
    has_sf = a.get("scaling_factor") is not None
    s = (a["scaling"].reshape(-1, 3).astype(np.float64) - float(a["scaling_zero_point"])) * float(a["scaling_scale"])
    if has_sf:
        s = np.maximum(s, 0.0); s = s / np.linalg.norm(s, axis=1, keepdims=True)
    else:
        s = np.exp(s)
    q = (a["rotation"].reshape(-1, 4).astype(np.float64) - float(a["rotation_zero_point"])) * float(a["rotation_scale"])
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                  np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                  np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], 1)
    L = R * s[:, None, :]
    M = L @ L.transpose(0, 2, 1)
    return np.stack([M[:, 0, 0], M[:, 0, 1], M[:, 0, 2], M[:, 1, 1], M[:, 1, 2], M[:, 2, 2]], -1)


@pytest.mark.parametrize("sf,idx,deg", [(True, True, 3), (False, True, 2), (True, False, 1), (False, False, 0)])
def test_oracle_c3dgs_convert(ws, orc, sf, idx, deg):
    n = 4000
    a = ws.synth.c3dgs_arrays(n, 11, deg, codebook=512, scaling_factor=sf, indices=idx)
    o = orc.c3dgs_convert(a)
    g = o["gaussians"].view(ws.synth.GAUSSIAN_COMPRESSED_DTYPE).reshape(-1)
    assert o["sh_deg"] == deg
    assert np.array_equal(g["xyz"], a["xyz"].astype(np.float32))
    assert np.array_equal(g["opacity"], a["opacity"].reshape(-1))
    assert np.array_equal(g["scale_factor"], a["scaling_factor"].reshape(-1) if sf else np.zeros(n, np.int8))
    assert np.array_equal(g["geometry_idx"], a["gaussian_indices"].astype(np.uint32) if idx else np.arange(n, dtype=np.uint32))
    assert np.array_equal(g["sh_idx"], a["feature_indices"].astype(np.uint32) if idx else np.arange(n, dtype=np.uint32))
    assert not g["_pad"].any()
    # SH codebook: dc then rest per entry (io/npz.rs:193-205)
    want = np.concatenate([a["features_dc"].reshape(len(a["features_dc"]), 3), a["features_rest"].reshape(len(a["features_dc"]), -1)], 1)
    assert np.array_equal(o["sh_coefs"], want)
    cov = o["covars"].view(np.uint16).reshape(-1, 6)
    d = np.abs(f16_ordered(cov) - f16_ordered(_np_covars(a).astype(np.float16).view(np.uint16)))
    assert d.max() <= 4 and (d > 0).mean() < 0.02
    # new_compressed grows the box from the unit cube (io/mod.rs:119)
    assert np.all(o["bbox"][:3] <= -1.0) and np.all(o["bbox"][3:] >= 1.0)
    assert np.allclose(o["center"], a["xyz"].astype(np.float64).mean(0), atol=1e-4)
    assert o["up"] is None


def test_synthetic_npz_round_trips_through_numpy(ws):
    a = ws.synth.c3dgs_arrays(100, 2, 2, codebook=16)
    import io
    with np.load(io.BytesIO(ws.synth.npz_bytes(a))) as z:
        assert set(z.files) == set(a)
        for k in a:
            assert np.array_equal(z[k], a[k])
        assert z["features_rest"].shape[1] + 1 == 9                      # io/npz.rs:33-37 derives sh_deg from this


@pytest.mark.gpu
@pytest.mark.parametrize("sf,idx,deg", [(True, True, 3), (False, True, 2), (True, False, 1), (False, False, 0)])
def test_gpu_npz_ingest_matches_oracle(ws, orc, ctx, sf, idx, deg):
    n = 30011
    a = ws.synth.c3dgs_arrays(n, 17, deg, codebook=2048, scaling_factor=sf, indices=idx)
    a["xyz"] = (a["xyz"].astype(np.float32) * 12.0).astype(np.float16)
    a["xyz"][:, 1] *= np.float16(0.05)                                   # big + flat: `up` is Some
    pc = ws.PointCloud.from_npz(ctx, ws.synth.npz_bytes(a))
    o = orc.c3dgs_convert(a)
    assert pc.num_points() == n and pc.sh_deg() == deg and pc.compressed()
    assert pc.mip_splatting() is True and abs(pc.dilation_kernel_size() - 0.1) < 1e-7 and np.array_equal(pc.background_color(), np.zeros(3))
    assert np.array_equal(pc.read("gaussians"), o["gaussians"])          # integer / copy work: bit-exact
    assert np.array_equal(pc.read("sh_coefs").reshape(o["sh_coefs"].shape), o["sh_coefs"])
    assert np.array_equal(pc.read("xyz"), a["xyz"].astype(np.float32))
    cov, ocov = pc.read("covars").view(np.uint16), o["covars"].view(np.uint16)
    d = np.abs(f16_ordered(cov) - f16_ordered(ocov))
    if sf:
        assert d.max() == 0                                              # no transcendental on this branch: bit-exact
    else:
        assert d.max() <= 1 and (d > 0).mean() < 2e-3                    # expf: CUDA vs glibc
    b = pc.bbox()
    assert np.array_equal(np.concatenate([b.min, b.max]).astype(np.float32), o["bbox"])
    assert np.allclose(pc.center(), o["center"], atol=1e-3)
    up = pc.up()
    assert up is not None and o["up"] is not None and np.dot(up, o["up"]) > 0.9999


@pytest.mark.gpu
def test_gpu_npz_cloud_renders_like_the_uploaded_cloud(ws, orc, ctx):
    n, W, H = 30000, 640, 360
    a = ws.synth.c3dgs_arrays(n, 5, 3, codebook=4096, metadata=False)      # metadata would change the resolved settings
    pc_npz = ws.PointCloud.from_npz(ctx, a)                               # already-decoded members
    o = orc.c3dgs_convert(a)
    zp = lambda k: (int(a[k + "_zero_point"]), np.float32(a[k + "_scale"]))
    sh = pc_npz.read("sh_coefs")
    sh = np.concatenate([sh, np.zeros((-sh.size) % 4 + 4, np.int8)])
    cloud = dict(gaussians=pc_npz.read("gaussians").view(ws.synth.GAUSSIAN_COMPRESSED_DTYPE).reshape(-1), sh_coefs=sh,
                 covars=pc_npz.read("covars").view(np.float16).reshape(-1, 6), num_points=n, sh_deg=3, compressed=True,
                 quantization=dict(color_dc=zp("features_dc"), color_rest=zp("features_rest"), opacity=zp("opacity"),
                                   scaling_factor=zp("scaling_factor")),
                 aabb_min=o["bbox"][:3], aabb_max=o["bbox"][3:], center=o["center"])
    pc_new = ws.PointCloud.new(ctx, make_generic(ws, cloud))
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    pos, rot = ws.synth.orbit_camera(120.0)
    args = make_args(ws, cloud, pos, rot, W, H, fovx, fovy)
    import torch
    frames = []
    for pc in (pc_npz, pc_new):
        r = ws.GaussianRenderer.new(ctx, ws.FORMAT_RGBA16_FLOAT, 3, True)
        r.prepare(None, pc, args)
        out = r.empty_host_frame()
        r.render_to_host(out, pc)
        torch.cuda.synchronize()
        frames.append(out.copy())
    assert np.array_equal(frames[0].view(np.uint16), frames[1].view(np.uint16))
    assert frames[0].view(np.uint16).any()
    # and the frame agrees with the oracle's render of the oracle-converted cloud (stage tolerances as in test_gpu_parity)
    ocloud = dict(cloud, gaussians=o["gaussians"].view(ws.synth.GAUSSIAN_COMPRESSED_DTYPE).reshape(-1),
                  covars=o["covars"].view(np.float16).reshape(-1, 6))
    ref = orc.render_frame(ocloud, pos, rot, W, H, fovx, fovy)
    img = frames[0].astype(np.float32)
    assert np.abs(img - ref["image"]).mean() < 1.5e-4              # f16 frame against the oracle's f32 image
