"""CPU tests of the host logic: synthetic layouts, camera conventions, uniform resolution."""
import math

import numpy as np


def test_orbit_camera_is_a_rotation_looking_at_origin(ws):
    for az in (0, 10, 90, 200, 350):
        eye, q = ws.synth.orbit_camera(az)
        R = ws.synth._quat_to_mat(q[None].astype(np.float64))[0]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and np.linalg.det(R) > 0.999
        assert abs(np.linalg.norm(eye) - 3.0) < 1e-5
        f = R[2]                                       # third row = forward axis
        assert np.allclose(f, -eye / np.linalg.norm(eye), atol=1e-6)
        # world->camera of the origin lies on the optical axis at distance 3
        assert np.allclose(R @ (np.zeros(3) - eye), (0, 0, 3), atol=1e-5)


def test_fov_gives_square_pixels(ws):
    for W, H in ((800, 600), (1200, 799), (1920, 1080), (3840, 2160)):
        fx, fy = ws.synth.fov_for_viewport(W, H)
        assert math.isclose(W / (2 * math.tan(fx / 2)), H / (2 * math.tan(fy / 2)), rel_tol=1e-9)


def test_build_cov_matches_reference_formula(ws):
    """utils.rs:194-204: M = (R S)(R S)^T, upper triangle."""
    rng = np.random.default_rng(1)
    q = rng.standard_normal((50, 4)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
    s = np.exp(rng.standard_normal((50, 3))).astype(np.float32)
    c = ws.synth.build_cov(q, s)
    R = ws.synth._quat_to_mat(q.astype(np.float64))
    M = R @ (np.eye(3) * (s.astype(np.float64) ** 2)[:, None, :]) @ R.transpose(0, 2, 1)
    ref = np.stack([M[:, 0, 0], M[:, 0, 1], M[:, 0, 2], M[:, 1, 1], M[:, 1, 2], M[:, 2, 2]], 1)
    assert np.allclose(c, ref, rtol=2e-5, atol=1e-7)


def test_synth_cloud_layout(ws):
    c = ws.synth.make_cloud(1000, 3)
    assert c["gaussians"].nbytes == 1000 * 28 and c["sh_coefs"].nbytes == 1000 * 96
    assert c["sh_coefs"].dtype == np.float16 and c["sh_coefs"].shape == (1000, 16, 3)
    assert (c["aabb_min"] <= 0).all() and (c["aabb_max"] >= 0).all()      # bbox grows from the zero box (io/mod.rs:74)
    op = c["gaussians"]["opacity"].astype(np.float32)
    assert (op >= 0).all() and (op <= 1).all()
    cc = ws.synth.make_cloud_compressed(1000, 3, codebook=64)
    assert cc["gaussians"].nbytes == 1000 * 24 and cc["covars"].shape == (64, 6)
    assert cc["gaussians"]["geometry_idx"].max() < 64 and cc["gaussians"]["sh_idx"].max() < 64
    assert cc["sh_coefs"].dtype == np.int8 and cc["sh_coefs"].size >= 64 * 48
    assert (cc["aabb_min"] <= -1).all() and (cc["aabb_max"] >= 1).all()  # starts from the unit cube (io/mod.rs:119)


def test_deterministic_generation(ws):
    a = ws.synth.make_cloud(500, 11); b = ws.synth.make_cloud(500, 11)
    assert a["gaussians"].tobytes() == b["gaussians"].tobytes() and a["sh_coefs"].tobytes() == b["sh_coefs"].tobytes()
