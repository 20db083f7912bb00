"""Synthetic Gaussian clouds and cameras in the reference's exact GPU byte layouts.

The benchmark inputs are synthetic: positions
uniform in [-1,1]^3, log-normal per-axis scales around 0.6*N^(-1/3), random rotations, sigmoid
opacities, SH dc ~ N(0,1) and higher coefficients ~ N(0,0.1^2).  Everything is rounded exactly as
the reference's loaders do it (opacity / covariance / SH to f16: io/ply.rs:95-98; covariance =
R S S^T R^T: utils.rs:194-204), so the buffers are what PointCloud::new would upload
(pointcloud.rs:119-129) and no loader is inside any timed region.
"""
import math

import numpy as np

GAUSSIAN_DTYPE = np.dtype([("xyz", "<f4", 3), ("opacity", "<f2"), ("_pad", "<f2"), ("cov", "<f2", 6)])   # 28 B, pointcloud.rs:38-45
GAUSSIAN_COMPRESSED_DTYPE = np.dtype([("xyz", "<f4", 3), ("opacity", "i1"), ("scale_factor", "i1"), ("_pad", "<u2"),
                                      ("geometry_idx", "<u4"), ("sh_idx", "<u4")])                       # 24 B, pointcloud.rs:14-24
assert GAUSSIAN_DTYPE.itemsize == 28 and GAUSSIAN_COMPRESSED_DTYPE.itemsize == 24

CONFIGS = {
    # name: (N, W, H, seed, compressed)      BASELINE.json configs / BASELINE.md section 3
    "cfg1": (100_000, 800, 600, 1001, False),
    "cfg2": (1_000_000, 1200, 799, 1002, False),
    "cfg3": (6_000_000, 1920, 1080, 1003, False),
    "cfg4": (6_000_000, 3840, 2160, 1004, True),
    "cfg5": (24_000_000, 1920, 1080, 1005, False),
}


def _quat_to_mat(q):
    """cgmath Matrix3::from(Quaternion) for (w,x,y,z) rows of q; returns math matrices [n,3,3] (row, col)."""
    s, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    x2, y2, z2 = x + x, y + y, z + z
    xx2, xy2, xz2 = x2 * x, x2 * y, x2 * z
    yy2, yz2, zz2 = y2 * y, y2 * z, z2 * z
    sy2, sz2, sx2 = y2 * s, z2 * s, x2 * s
    m = np.empty((q.shape[0], 3, 3), dtype=q.dtype)
    # cgmath columns: c0 = (1-yy2-zz2, xy2+sz2, xz2-sy2), c1 = (xy2-sz2, 1-xx2-zz2, yz2+sx2), c2 = (xz2+sy2, yz2-sx2, 1-xx2-yy2)
    m[:, 0, 0] = 1 - yy2 - zz2; m[:, 1, 0] = xy2 + sz2; m[:, 2, 0] = xz2 - sy2
    m[:, 0, 1] = xy2 - sz2; m[:, 1, 1] = 1 - xx2 - zz2; m[:, 2, 1] = yz2 + sx2
    m[:, 0, 2] = xz2 + sy2; m[:, 1, 2] = yz2 - sx2; m[:, 2, 2] = 1 - xx2 - yy2
    return m


def build_cov(rot_wxyz, scale):
    """utils.rs:194-204: upper triangle (xx,xy,xz,yy,yz,zz) of R S S^T R^T, f32."""
    R = _quat_to_mat(rot_wxyz.astype(np.float32))
    L = R * scale.astype(np.float32)[:, None, :]
    M = np.einsum("nij,nkj->nik", L, L)
    return np.stack([M[:, 0, 0], M[:, 0, 1], M[:, 0, 2], M[:, 1, 1], M[:, 1, 2], M[:, 2, 2]], axis=1).astype(np.float32)


def _attributes(n, seed, chunk=None, density_n=None):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1.0, 1.0, size=(n, 3)).astype(np.float32)
    s0 = 0.6 * max(density_n or n, 1) ** (-1.0 / 3.0)        # splat size follows the density of the WHOLE cloud
    scale = (s0 * np.exp(0.6 * rng.standard_normal((n, 3)))).astype(np.float32)
    q = rng.standard_normal((n, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opacity = (1.0 / (1.0 + np.exp(-2.0 * rng.standard_normal(n)))).astype(np.float32)
    sh = rng.standard_normal((n, 16, 3), dtype=np.float32)     # f32 generator: 6M x 48 values is the bulk of the time
    sh[:, 1:, :] *= np.float32(0.1)
    return xyz, scale, q, opacity, sh


def _bbox_center(xyz, compressed):
    # io/mod.rs:74,119: raw clouds start the bbox from the zero box, compressed from the unit cube
    lo = xyz.min(axis=0) if len(xyz) else np.zeros(3, np.float32)
    hi = xyz.max(axis=0) if len(xyz) else np.zeros(3, np.float32)
    if compressed:
        lo = np.minimum(lo, -1.0); hi = np.maximum(hi, 1.0)
    else:
        lo = np.minimum(lo, 0.0); hi = np.maximum(hi, 0.0)
    center = xyz.mean(axis=0, dtype=np.float64).astype(np.float32) if len(xyz) else np.zeros(3, np.float32)
    return lo.astype(np.float32), hi.astype(np.float32), center


def make_cloud(n, seed, sh_deg=3, density_n=None):
    """Raw layout (28-B Gaussian + 96-B SH).  Returns a dict of numpy buffers + metadata.
    `density_n`: total size of the cloud this one is a part of (a shard generated on its own keeps the splat size of the whole)."""
    xyz, scale, q, opacity, sh = _attributes(n, seed, density_n=density_n)
    g = np.zeros(n, dtype=GAUSSIAN_DTYPE)
    g["xyz"] = xyz
    g["opacity"] = opacity.astype(np.float16)
    g["cov"] = build_cov(q, scale).astype(np.float16)
    sh16 = sh.astype(np.float16)                       # [[f16;3];16]
    lo, hi, center = _bbox_center(xyz, False)
    return dict(gaussians=g, sh_coefs=sh16, num_points=n, sh_deg=sh_deg, compressed=False,
                aabb_min=lo, aabb_max=hi, center=center)


def make_cloud_compressed(n, seed, sh_deg=3, codebook=4096, identity_index=False):
    """npz/c3dgs layout (io/npz.rs:59-225): 24-B records, f16 covariance codebook of the
    scale-normalised covariance, i8 SH codebook, i8 opacity and i8 log scale factor."""
    xyz, scale, q, opacity, sh = _attributes(n, seed)
    rng = np.random.default_rng(seed + 77)
    k_geo = n if identity_index else min(codebook, max(n, 1))
    k_sh = n if identity_index else min(codebook, max(n, 1))
    # geometry: covariance of the unit-norm scale, plus a per-splat log scale factor
    norm = np.linalg.norm(scale, axis=1)
    if identity_index:
        cov_cb = build_cov(q, scale / norm[:, None]).astype(np.float16)
        geo_idx = np.arange(n, dtype=np.uint32)
        sh_cb = sh
        sh_idx = np.arange(n, dtype=np.uint32)
    else:
        pick = rng.integers(0, n, size=k_geo) if n else np.zeros(0, np.int64)
        cov_cb = build_cov(q[pick], scale[pick] / norm[pick][:, None]).astype(np.float16)
        geo_idx = rng.integers(0, k_geo, size=n).astype(np.uint32)
        pick_s = rng.integers(0, n, size=k_sh) if n else np.zeros(0, np.int64)
        sh_cb = sh[pick_s]
        sh_idx = rng.integers(0, k_sh, size=n).astype(np.uint32)
    ncoef = (sh_deg + 1) ** 2

    def quant(v, lo, hi):
        sc = (hi - lo) / 255.0
        zp = int(round(-128 - lo / sc))
        qv = np.clip(np.round(v / sc + zp), -128, 127).astype(np.int8)
        return qv, zp, np.float32(sc)

    logs = np.log(norm).astype(np.float32)
    q_sf, zp_sf, sc_sf = quant(logs, float(logs.min()) if n else -1.0, float(logs.max()) if n else 1.0)
    q_op, zp_op, sc_op = quant(opacity, 0.0, 1.0)
    dc = sh_cb[:, 0, :]
    rest = sh_cb[:, 1:ncoef, :]
    q_dc, zp_dc, sc_dc = quant(dc, -4.0, 4.0)
    q_rest, zp_rest, sc_rest = quant(rest, -0.5, 0.5)
    sh_i8 = np.concatenate([q_dc[:, None, :], q_rest], axis=1).reshape(len(sh_cb), ncoef * 3)   # dc first, then rest
    sh_bytes = np.ascontiguousarray(sh_i8).reshape(-1)
    pad = (-sh_bytes.size) % 4 + 4                         # the shader reads whole u32 words (+1 look-ahead)
    sh_bytes = np.concatenate([sh_bytes, np.zeros(pad, np.int8)])

    g = np.zeros(n, dtype=GAUSSIAN_COMPRESSED_DTYPE)
    g["xyz"] = xyz.astype(np.float16).astype(np.float32)   # f16 positions in the file, io/npz.rs:96-100
    g["opacity"] = q_op
    g["scale_factor"] = q_sf
    g["geometry_idx"] = geo_idx
    g["sh_idx"] = sh_idx
    lo, hi, center = _bbox_center(g["xyz"], True)
    quantization = dict(color_dc=(zp_dc, sc_dc), color_rest=(zp_rest, sc_rest),
                        opacity=(zp_op, sc_op), scaling_factor=(zp_sf, sc_sf))
    return dict(gaussians=g, sh_coefs=sh_bytes, covars=cov_cb, num_points=n, sh_deg=sh_deg, compressed=True,
                quantization=quantization, aabb_min=lo, aabb_max=hi, center=center)


# ---- cameras ---------------------------------------------------------------------------------
def _mat_to_quat_wxyz(R):
    """Unit quaternion (w,x,y,z) with cgmath Matrix3::from(q) == R (R = math matrix, row/col), f64 -> f32."""
    m = np.asarray(R, dtype=np.float64)
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        w, x, y, z = 0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        w, x, y, z = (m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s
    elif m[1, 1] > m[2, 2]:
        s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        w, x, y, z = (m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s
    else:
        s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        w, x, y, z = (m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s
    q = np.array([w, x, y, z], dtype=np.float64)
    return (q / np.linalg.norm(q)).astype(np.float32)


def fov_for_viewport(width, height, fovy_deg=40.0):
    """Square pixels: fx == fy."""
    fovy = math.radians(fovy_deg)
    fovx = 2.0 * math.atan(math.tan(fovy / 2.0) * width / height)
    return fovx, fovy


def fixed_camera():
    """cfg 1: eye (0,0,-3), identity rotation (camera looks down +z, +y down: camera.rs:59-73)."""
    return np.array([0.0, 0.0, -3.0], np.float32), np.array([1.0, 0.0, 0.0, 0.0], np.float32)


def orbit_camera(az_deg, radius=3.0, elev_deg=15.0):
    """cfg 2-5 orbit: eye = r*(cos e sin az, -sin e, -cos e cos az); world->camera rows
    (r, d, f) with f = -eye/|eye|, r = normalize((0,1,0) x f), d = f x r."""
    az, el = math.radians(az_deg), math.radians(elev_deg)
    eye = radius * np.array([math.cos(el) * math.sin(az), -math.sin(el), -math.cos(el) * math.cos(az)])
    f = -eye / np.linalg.norm(eye)
    r = np.cross(np.array([0.0, 1.0, 0.0]), f)
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    R = np.stack([r, d, f], axis=0)          # rows
    return eye.astype(np.float32), _mat_to_quat_wxyz(R)


def orbit_views(count=36):
    return [orbit_camera(360.0 * i / count) for i in range(count)]


# ---- synthetic .ply files (the 3DGS training-output format io/ply.rs reads) ---------------------------
def ply_vertices(n, seed, sh_deg=3, spread=1.0):
    """(n, 14 + 3*(sh_deg+1)^2) float32 vertex block in the property order read_line assumes
    (io/ply.rs:50-100): x y z nx ny nz f_dc[3] f_rest[3][C-1] opacity(logit) scale(log)[3] rot(wxyz, unnormalised)[4]."""
    xyz, scale, q, opacity, sh = _attributes(n, seed)
    rng = np.random.default_rng(seed + 1234)
    C_ = (sh_deg + 1) ** 2
    v = np.zeros((n, 14 + 3 * C_), np.float32)
    v[:, 0:3] = xyz * np.float32(spread)
    v[:, 3:6] = rng.standard_normal((n, 3)).astype(np.float32)          # normals: present in the file, never used
    v[:, 6:9] = sh[:, 0, :]
    v[:, 9:9 + 3 * (C_ - 1)] = sh[:, 1:C_, :].transpose(0, 2, 1).reshape(n, -1)   # channel-major [3][C-1]
    k = 9 + 3 * (C_ - 1)
    with np.errstate(divide="ignore"):
        v[:, k] = np.log(opacity / (1.0 - opacity))
    v[:, k + 1:k + 4] = np.log(scale * np.float32(spread))
    v[:, k + 4:k + 8] = q * rng.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)   # training output is not unit length
    return v


def ply_bytes(vertices, sh_deg, comments=(), big_endian=False, fmt=None):
    """Serialise a vertex block into a .ply file image (header + binary body)."""
    C_ = (sh_deg + 1) ** 2
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    names += ["f_rest_%d" % i for i in range(3 * (C_ - 1))]
    names += ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert vertices.shape[1] == len(names)
    fmt = fmt or ("binary_big_endian" if big_endian else "binary_little_endian")
    head = ["ply", "format %s 1.0" % fmt] + ["comment %s" % c for c in comments]
    head += ["element vertex %d" % len(vertices)] + ["property float %s" % nm for nm in names] + ["end_header"]
    body = np.ascontiguousarray(vertices, dtype=">f4" if big_endian else "<f4").tobytes()
    return ("\n".join(head) + "\n").encode("ascii") + body


# ---- synthetic compressed .npz members (what io/npz.rs reads) -----------------------------------------
def c3dgs_arrays(n, seed, sh_deg=3, codebook=4096, scaling_factor=True, indices=True, metadata=True):
    """dict of the arrays a c3dgs .npz holds (io/npz.rs:58-160).  scaling_factor=False gives the older variant
    whose `scaling` is the quantised log-scale; indices=False the variant with one codebook entry per point."""
    xyz, scale, q, opacity, sh = _attributes(n, seed)
    rng = np.random.default_rng(seed + 99)
    k = min(codebook, max(n, 1)) if indices else n
    pick_g = rng.integers(0, max(n, 1), size=k) if (indices and n) else np.arange(k)
    pick_s = rng.integers(0, max(n, 1), size=k) if (indices and n) else np.arange(k)
    C_ = (sh_deg + 1) ** 2

    def quant(v, lo, hi):
        sc = (hi - lo) / 255.0
        zp = int(round(-128 - lo / sc))
        return np.clip(np.round(np.asarray(v, np.float64) / sc + zp), -128, 127).astype(np.int8), zp, np.float32(sc)

    out = {"xyz": xyz.astype(np.float16)}
    out["opacity"], out["opacity_zero_point"], out["opacity_scale"] = quant(opacity[:, None], 0.0, 1.0)
    norm = np.linalg.norm(scale, axis=1)
    if scaling_factor:
        logs = np.log(norm)
        out["scaling_factor"], out["scaling_factor_zero_point"], out["scaling_factor_scale"] = quant(
            logs[:, None], float(logs.min()) if n else -1.0, float(logs.max()) if n else 1.0)
        out["scaling"], out["scaling_zero_point"], out["scaling_scale"] = quant((scale / norm[:, None])[pick_g], -0.1, 1.0)
    else:
        ls = np.log(scale[pick_g])
        out["scaling"], out["scaling_zero_point"], out["scaling_scale"] = quant(ls, float(ls.min()) if n else -1.0, float(ls.max()) if n else 1.0)
    out["rotation"], out["rotation_zero_point"], out["rotation_scale"] = quant(q[pick_g], -1.0, 1.0)
    out["features_dc"], out["features_dc_zero_point"], out["features_dc_scale"] = quant(sh[pick_s][:, 0:1, :], -4.0, 4.0)
    out["features_rest"], out["features_rest_zero_point"], out["features_rest_scale"] = quant(sh[pick_s][:, 1:C_, :], -0.5, 0.5)
    if indices:
        out["gaussian_indices"] = rng.integers(0, k, size=n).astype(np.int32)
        out["feature_indices"] = rng.integers(0, k, size=n).astype(np.int32)
    if metadata:
        out["kernel_size"] = np.float32(0.1)
        out["mip_splatting"] = np.bool_(True)
        out["background_color"] = np.array([0.0, 0.0, 0.0], np.float32)
    return out


def npz_bytes(arrays):
    """Serialise the members into a .npz file image."""
    import io
    f = io.BytesIO()
    np.savez(f, **arrays)
    return f.getvalue()
