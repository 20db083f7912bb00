"""CPU tests of the oracle itself: pinned against the reference's only KAT, cross-checked against an
independent numpy restatement and closed-form cases, and frozen by golden fixtures."""
import math
import os

import numpy as np
import pytest

from helpers import f16_ordered

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_f16_conversion_matches_ieee(orc):
    # every half round-trips; random floats round like numpy's RNE conversion
    allh = np.arange(65536, dtype=np.uint16)
    f = np.array([orc.f16_bits_to_f32(h) for h in allh[::7]], np.float32)
    ref = allh[::7].view(np.float16).astype(np.float32)
    assert np.array_equal(np.isnan(f), np.isnan(ref))
    assert np.array_equal(f[~np.isnan(f)].view(np.uint32), ref[~np.isnan(ref)].view(np.uint32))
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(4000).astype(np.float32) * s for s in (1e-8, 1e-5, 1e-3, 1, 300, 70000)])
    x = np.concatenate([x, np.array([0, -0.0, 65504, 65519.99, 65520, 1e9, -1e9, 5.96e-8, 2.98e-8, 2.99e-8, 6.1e-5], np.float32)])
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    got = np.array([orc.f32_to_f16_bits(v) for v in x], np.uint16)
    assert np.array_equal(got, want)


def test_sort_kat_from_reference(orc):
    """GPURSSorter::test_sort, gpu_rs.rs:295-331: f32 keys 8191.0 .. 0.0 must come out 0.0 .. 8191.0."""
    n = 8192
    keys = np.arange(n - 1, -1, -1, dtype=np.float32).view(np.uint32)
    k, v = orc.sort_pairs(keys, np.arange(n, dtype=np.uint32))
    assert np.array_equal(k.view(np.float32), np.arange(n, dtype=np.float32))
    assert np.array_equal(v, np.arange(n - 1, -1, -1, dtype=np.uint32))


def test_sort_stable_against_numpy(orc):
    rng = np.random.default_rng(3)
    for n in (0, 1, 255, 4097, 50000):
        keys = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
        keys[: n // 3] &= 0xff            # many ties -> stability is exercised
        k, v = orc.sort_pairs(keys, np.arange(n, dtype=np.uint32))
        o = np.argsort(keys, kind="stable")
        assert np.array_equal(k, keys[o]) and np.array_equal(v, o.astype(np.uint32))


def test_sh_constants_match_reference():
    """preprocess.wgsl:4-23."""
    from oracle import np_oracle
    assert float(np_oracle.SH_C0) == pytest.approx(0.28209479177387814, rel=1e-7)
    assert float(np_oracle.SH_C1) == pytest.approx(0.4886025119029199, rel=1e-7)
    assert np.allclose(np_oracle.SH_C2, [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396])
    assert np.allclose(np_oracle.SH_C3[[0, 1, 5]], [-0.5900435899266435, 2.890611442640554, 1.445305721320277])


@pytest.mark.parametrize("az", [0.0, 70.0, 200.0])
def test_stage1_c_vs_numpy(orc, ws, az):
    from oracle import np_oracle
    cloud = ws.synth.make_cloud(20000, 7)
    pos, rot = ws.synth.orbit_camera(az)
    W, H = 640, 360
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    zn, zf = orc.fit_near_far(pos, cloud["aabb_min"], cloud["aabb_max"])
    cam = orc.camera_uniform(pos, rot, fovx, fovy, zn, zf, W, H)
    st = orc.render_settings(cloud)
    s, k, src = orc.preprocess(cloud, cam, st)
    keep, s2, k2 = np_oracle.preprocess_raw(cloud, cam, st)
    assert np.array_equal(np.nonzero(keep)[0], src)
    a = s.view(np.float16).astype(np.float64); b = s2.astype(np.float64)
    # centre, colour, opacity: 1 f16 ulp
    d = np.abs(f16_ordered(s[:, 4:]) - f16_ordered(s2.view(np.uint16)[:, 4:]))
    assert d.max() <= 1
    # axes: the eigenvector formula cancels for near-axis-aligned splats (normalize((off, l1-d1)),
    # preprocess.wgsl:248), so compare each axis as a vector, relative to its length
    for sl in (slice(0, 2), slice(2, 4)):
        na = np.linalg.norm(a[:, sl] * [W, H], axis=1)
        err = np.linalg.norm((a[:, sl] - b[:, sl]) * [W, H], axis=1)
        assert (err <= 1e-2 * na + 1e-6).all() and (err <= 2e-3 * na + 1e-6).mean() > 0.999
    assert np.abs(k.astype(np.int64) - k2.astype(np.int64)).max() <= 64


def _one_gaussian_cloud(ws, xyz, sigma, opacity, rgb):
    g = np.zeros(1, dtype=ws.synth.GAUSSIAN_DTYPE)
    g["xyz"] = np.asarray(xyz, np.float32)
    g["opacity"] = np.float16(opacity)
    g["cov"] = np.array([sigma * sigma, 0, 0, sigma * sigma, 0, sigma * sigma], np.float16)
    sh = np.zeros((1, 16, 3), np.float16)
    sh[0, 0] = (np.asarray(rgb, np.float64) - 0.5) / 0.28209479177387814
    return dict(gaussians=g, sh_coefs=sh, num_points=1, sh_deg=3, compressed=False,
                aabb_min=np.array([-1, -1, -1], np.float32), aabb_max=np.array([1, 1, 1], np.float32),
                center=np.zeros(3, np.float32))


def test_analytic_isotropic_gaussian_on_axis(orc, ws):
    """Closed form: an isotropic Gaussian on the optical axis projects to an isotropic 2D Gaussian
    with variance (f*sigma/z)^2 + kernel_size; alpha(r) = min(.99, o*exp(-r^2/(2 var))) inside
    r^2/(2 var) <= 2*CUTOFF, 0 outside (gaussian.wgsl:59-66)."""
    W = H = 257
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    sigma, o = 0.02, 0.8
    # on-axis would hit normalize((0,0)); nudge it off the axis so off-diagonal != 0
    cloud = _one_gaussian_cloud(ws, (0.013, 0.007, 0.0), sigma, o, (0.9, 0.4, 0.2))
    pos, rot = ws.synth.fixed_camera()
    fr = orc.render_frame(cloud, pos, rot, W, H, fovx, fovy)
    assert len(fr["keys"]) == 1
    img = fr["image"]
    f = H / (2 * math.tan(fovy / 2))
    z = 3.0
    var = (f * sigma / z) ** 2 + 0.3
    cxp = W / 2 + f * 0.013 / z; cyp = H / 2 + f * 0.007 / z     # +y is down in camera space and in the image
    ys, xs = np.mgrid[0:H, 0:W]
    r2 = (xs + 0.5 - cxp) ** 2 + (ys + 0.5 - cyp) ** 2
    a = r2 / (2 * var)
    alpha = np.where(a <= 2 * 2.3539888583335364, np.minimum(0.99, o * np.exp(-a)), 0.0)
    # f16 storage of axes/centre: a few 1e-3 relative on the footprint size
    band = np.abs(a - 2 * 2.3539888583335364) < 0.15
    assert np.abs(img[..., 3] - alpha)[~band].max() < 6e-3
    assert np.abs(img[..., 0] - 0.9 * alpha)[~band].max() < 6e-3
    assert img[..., 3].max() > 0.75


def test_composite_order_and_clear(orc, ws):
    """Two overlapping splats: 'over' is order dependent, back-to-front = ascending key; the clear
    colour shows through with weight prod(1-b) (renderer.rs:63-67, lib.rs:451-462)."""
    W = H = 64
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    c1 = _one_gaussian_cloud(ws, (0.01, 0.005, 0.0), 0.15, 0.9, (1.0, 0.0, 0.0))
    c2 = _one_gaussian_cloud(ws, (0.012, 0.004, 0.5), 0.15, 0.9, (0.0, 0.0, 1.0))
    g = np.concatenate([c1["gaussians"], c2["gaussians"]]); sh = np.concatenate([c1["sh_coefs"], c2["sh_coefs"]])
    cloud = dict(c1, gaussians=g, sh_coefs=sh, num_points=2)
    pos, rot = ws.synth.fixed_camera()
    fr = orc.render_frame(cloud, pos, rot, W, H, fovx, fovy, clear=(0, 1, 0, 1))
    # splat 0 (z=0) is nearer than splat 1 (z=0.5): larger key = nearer => drawn last
    assert fr["keys"][0] > fr["keys"][1] and list(fr["order"]) == [1, 0]
    px = fr["image"][H // 2, W // 2]
    assert px[0] > 0.85 and px[2] < 0.12 and px[1] < 0.02       # red on top, green clear mostly hidden
    assert px[3] == pytest.approx(1.0, abs=1e-6)
    corner = fr["image"][0, 0]
    assert np.allclose(corner, (0, 1, 0, 1))                    # untouched pixel = clear colour


def test_culling_rules(orc, ws):
    """preprocess.wgsl:177,190: clip box, 0<z<1, 1.2 frustum margin."""
    W, H = 320, 200
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    pos, rot = ws.synth.fixed_camera()
    pts = [(0.01, 0.01, 0.0), (0.01, 0.01, -4.0), (30.0, 0.0, 0.0), (0.0, 0.0, 0.5), (1.5, 0.0, 0.0)]
    g = np.zeros(len(pts), dtype=ws.synth.GAUSSIAN_DTYPE)
    g["xyz"] = np.asarray(pts, np.float32); g["opacity"] = np.float16(0.5)
    g["cov"] = np.array([1e-3, 1e-5, 0, 1e-3, 0, 1e-3], np.float16)
    cloud = dict(gaussians=g, sh_coefs=np.zeros((len(pts), 16, 3), np.float16), num_points=len(pts), sh_deg=3,
                 compressed=False, aabb_min=np.array([-1, -1, -1], np.float32), aabb_max=np.array([1, 1, 1], np.float32),
                 center=np.zeros(3, np.float32))
    fr = orc.render_frame(cloud, pos, rot, W, H, fovx, fovy)
    zn, zf = orc.fit_near_far(pos, cloud["aabb_min"], cloud["aabb_max"])
    cam = orc.camera_uniform(pos, rot, fovx, fovy, zn, zf, W, H)
    _, _, src = orc.preprocess(cloud, cam, orc.render_settings(cloud))
    # 1: behind the camera; 2 and 4: outside the clip box (= bbox [-1,1]^3)
    assert list(src) == [0, 3]


def test_golden_fixture_frozen(orc, ws):
    """tests/golden/oracle_small.npz was produced by tests/golden/make_golden.py from this oracle;
    it freezes the restatement (any later edit of ws_oracle.c that changes results must be deliberate)."""
    z = np.load(os.path.join(GOLDEN, "oracle_small.npz"))
    cloud = ws.synth.make_cloud(int(z["n"]), int(z["seed"]))
    pos, rot = ws.synth.orbit_camera(float(z["az"]))
    W, H = int(z["W"]), int(z["H"])
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    fr = orc.render_frame(cloud, pos, rot, W, H, fovx, fovy)
    assert np.array_equal(fr["splats"], z["splats"])
    assert np.array_equal(fr["keys"], z["keys"])
    assert np.array_equal(fr["order"], z["order"])
    assert np.allclose(fr["image"][::4, ::4], z["image_sub"], atol=1e-6)
    assert orc.tile_rects(fr["splats"], W, H)[1] == int(z["pairs"])


# ---- oracle hardening (round 2): an independent float64 renderer and the ROP-faithful compositor ----------------
@pytest.mark.parametrize("az", [30.0, 200.0])
def test_f64_ideal_renderer_agrees_with_the_f16_faithful_oracle(orc, ws, az):
    """oracle/f64_ideal.py renders from the textbook formulation (pinhole Jacobian, 2x2 covariance, Mahalanobis
    footprint, no f16 intermediate, no eigen-axis quad) in float64; ws_oracle.c follows the shaders line by line.
    They may differ only by what the f16 pack of axes / centre / colour legitimately costs (measured: max 8e-3,
    mean 2.6e-4 on these scenes).  A misread transpose, sign, cutoff or blend order is orders of magnitude larger."""
    from oracle import f64_ideal
    n, W, H = 3000, 160, 120
    cloud = ws.synth.make_cloud(n, 99)
    pos, rot = ws.synth.orbit_camera(az)
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    clear = (0.1, 0.2, 0.3, 1.0)
    fr = orc.render_frame(cloud, pos, rot, W, H, fovx, fovy, clear=clear)
    zn, zf = orc.fit_near_far(pos, cloud["aabb_min"], cloud["aabb_max"])
    img, idx = f64_ideal.render(cloud, pos, rot, W, H, fovx, fovy, zn, zf, clear=clear)
    assert len(idx) == len(fr["keys"])
    d = np.abs(img - fr["image"])
    assert d.max() < 2e-2 and d.mean() < 1e-3, (d.max(), d.mean())
    # and the scene is not trivially empty / saturated
    assert 0.2 < fr["image"][..., 3].mean() <= 1.0 and np.abs(fr["image"][..., :3] - np.array(clear[:3])).max() > 0.3


def test_rop_faithful_compositor(orc, ws):
    """wso_composite_rop rounds the destination to the target format after EVERY blend, as the reference's render
    target does (renderer.rs:63-67).  f32 target == the float oracle; one layer on a clear == one rounding; the gap of
    a float compositor (one conversion at the end) stays inside the bounds DESIGN.md quotes (profiles/r02_rop_gap.json)."""
    n, W, H = 20000, 320, 200
    cloud = ws.synth.make_cloud(n, 5)
    pos, rot = ws.synth.orbit_camera(75.0)
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    fr = orc.render_frame(cloud, pos, rot, W, H, fovx, fovy)
    assert np.array_equal(orc.composite_rop(fr["splats"], fr["order"], W, H, 2), fr["image"])
    r16 = orc.composite_rop(fr["splats"], fr["order"], W, H, 1)
    assert np.array_equal(r16, r16.astype(np.float16).astype(np.float32))            # every value is an f16
    flt16 = fr["image"].astype(np.float16).astype(np.float32)
    rel = np.abs(flt16 - r16).max(axis=2) / np.maximum(np.abs(r16).max(axis=2), 2.0 ** -10)
    assert rel.max() < 6e-3 and rel.mean() < 1e-3
    r8 = orc.composite_rop(fr["splats"], fr["order"], W, H, 0)
    assert np.array_equal(np.rint(r8 * 255.0), r8 * 255.0) or np.abs(np.rint(r8 * 255.0) - r8 * 255.0).max() < 1e-4
    assert r8.min() >= 0.0 and r8.max() <= 1.0
    # one splat over a clear colour: exactly one rounding of the blend
    one = _one_gaussian_cloud(ws, (0.01, 0.005, 0.0), 0.15, 0.9, (0.8, 0.3, 0.1))
    f1 = orc.render_frame(one, *ws.synth.fixed_camera(), 64, 64, *ws.synth.fov_for_viewport(64, 64), clear=(0.25, 0.5, 0.75, 1.0))
    q8 = orc.composite_rop(f1["splats"], f1["order"], 64, 64, 0, clear=(0.25, 0.5, 0.75, 1.0))
    c8 = np.rint(np.array([0.25, 0.5, 0.75, 1.0]) * 255.0) / 255.0                   # the clear is stored in the target first
    img1 = orc.composite(f1["splats"], f1["order"], 64, 64, clear=c8.astype(np.float32))
    assert np.abs(q8 - np.rint(np.clip(img1, 0, 1) * 255.0) / 255.0).max() < 1e-6
    # LDR scene (colours <= 1): the float compositor is within a few 8-bit steps of the per-layer-rounded target
    sp = fr["splats"].copy()
    c = sp[:, 6:9].view(np.float16); c[:] = np.minimum(c, np.float16(1.0))
    img_l = orc.composite(sp, fr["order"], W, H)
    gap = np.abs(np.rint(np.clip(img_l, 0, 1) * 255.0) / 255.0 - orc.composite_rop(sp, fr["order"], W, H, 0)).max(axis=2)
    assert gap.max() <= 8.0 / 255.0 + 1e-6 and gap.mean() <= 1.5 / 255.0
