"""Second, independent restatement of stage 1 (preprocess.wgsl:163-280) in vectorised numpy.

TEST INFRASTRUCTURE ONLY.  Written matrix-style straight from the WGSL (mat3x3 products via
einsum) rather than as the scalar expansion in ws_oracle.c, so that a misreading of the shader's
column-major conventions in one of the two would show up as a disagreement
(tests/test_oracle.py::test_stage1_c_vs_numpy).  float32 throughout, np.float16 (RNE) for the
f16 pack.  Summation order differs from the C oracle, so agreement is to 1 f16 ulp / a few f32
ulp, not bit-exact.
"""
import numpy as np

SH_C0 = np.float32(0.28209479177387814)
SH_C1 = np.float32(0.4886025119029199)
SH_C2 = np.array([1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
                  -1.0925484305920792, 0.5462742152960396], np.float32)
SH_C3 = np.array([-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
                  -0.4570457994644658, 1.445305721320277, -0.5900435899266435], np.float32)


def _mat4(cols16):
    """column-major 16 floats -> math matrix [row, col]."""
    return np.asarray(cols16, np.float32).reshape(4, 4).T.copy()


def evaluate_sh(d, sh, deg):
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res + (-SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3])
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = res + (SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2.0 * zz - xx - yy) * sh[:, 6]
                         + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = res + (SH_C3[0] * y * (3.0 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
                             + SH_C3[2] * y * (4.0 * zz - xx - yy) * sh[:, 11]
                             + SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * sh[:, 12]
                             + SH_C3[4] * x * (4.0 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
                             + SH_C3[6] * x * (xx - 3.0 * yy) * sh[:, 15])
    return (res + np.float32(0.5)).astype(np.float32)


def preprocess_raw(cloud, cam, st):
    """cam: oracle.CameraUniform, st: oracle.RenderSettings.  Returns (visible mask [N],
    splats f16 [V,10], keys u32 [V]) in Gaussian index order."""
    g = cloud["gaussians"]
    xyz = g["xyz"].astype(np.float32)
    n = len(xyz)
    view, proj = _mat4(cam.view[:]), _mat4(cam.proj[:])
    view_inv = _mat4(cam.view_inv[:])
    clip_lo = np.array(st.clip_min[:3], np.float32); clip_hi = np.array(st.clip_max[:3], np.float32)
    keep = ~((xyz < clip_lo).any(1) | (xyz > clip_hi).any(1))
    p4 = np.concatenate([xyz, np.ones((n, 1), np.float32)], 1)
    camspace = (p4 @ view.T).astype(np.float32)
    pos2d = (camspace @ proj.T).astype(np.float32)
    with np.errstate(all="ignore"):
        bounds = np.float32(1.2) * pos2d[:, 3]
        z = pos2d[:, 2] / pos2d[:, 3]
        keep &= ~((z <= 0) | (z >= 1) | (pos2d[:, 0] < -bounds) | (pos2d[:, 0] > bounds) | (pos2d[:, 1] < -bounds) | (pos2d[:, 1] > bounds))
    idx = np.nonzero(keep)[0]
    xyz, camspace, pos2d = xyz[idx], camspace[idx], pos2d[idx]
    cov = g["cov"][idx].astype(np.float32)
    opacity = g["opacity"][idx].astype(np.float32)
    sh = cloud["sh_coefs"][idx].astype(np.float32)            # [V,16,3]

    center = np.array(st.center[:3], np.float32)
    dd = np.float32(5.0) * np.linalg.norm(center - xyz, axis=1).astype(np.float32) / np.float32(st.scene_extend)
    t = np.clip(np.float32(st.walltime) - dd, 0, 1).astype(np.float32)
    scale_mod = np.where(np.float32(st.walltime) > dd, t * t * (3 - 2 * t), 0).astype(np.float32)
    scaling = np.float32(st.gaussian_scaling) * scale_mod

    V = len(idx)
    Vrk = np.empty((V, 3, 3), np.float32)                     # symmetric: row/col irrelevant
    Vrk[:, 0, 0], Vrk[:, 0, 1], Vrk[:, 0, 2] = cov[:, 0], cov[:, 1], cov[:, 2]
    Vrk[:, 1, 0], Vrk[:, 1, 1], Vrk[:, 1, 2] = cov[:, 1], cov[:, 3], cov[:, 4]
    Vrk[:, 2, 0], Vrk[:, 2, 1], Vrk[:, 2, 2] = cov[:, 2], cov[:, 4], cov[:, 5]
    Vrk = Vrk * (scaling * scaling)[:, None, None]

    fx, fy = np.float32(cam.focal[0]), np.float32(cam.focal[1])
    cx, cy, cz = camspace[:, 0], camspace[:, 1], camspace[:, 2]
    # WGSL J given by columns; as a math matrix J[row, col] = J_wgsl[col][row]
    J = np.zeros((V, 3, 3), np.float32)
    J[:, 0, 0] = fx / cz;                J[:, 1, 0] = 0;                 J[:, 2, 0] = -(fx * cx) / (cz * cz)   # column 0
    J[:, 0, 1] = 0;                      J[:, 1, 1] = -fy / cz;          J[:, 2, 1] = (fy * cy) / (cz * cz)    # column 1
    Wm = view[:3, :3].T.copy()                                 # transpose(mat3(view[0].xyz, view[1].xyz, view[2].xyz))
    T = np.einsum("ij,njk->nik", Wm, J).astype(np.float32)
    cov2 = np.einsum("nji,njk,nkl->nil", T, Vrk, T).astype(np.float32)    # T^T V T
    # WGSL cov[c][r] = math cov2[r, c]
    c00, c01, c11 = cov2[:, 0, 0], cov2[:, 1, 0], cov2[:, 1, 1]

    ks = np.float32(st.kernel_size)
    if st.mip_splatting:
        det0 = np.maximum(np.float32(1e-6), c00 * c11 - c01 * c01)
        det1 = np.maximum(np.float32(1e-6), (c00 + ks) * (c11 + ks) - c01 * c01)
        coef = np.sqrt(det0 / (det1 + np.float32(1e-6)) + np.float32(1e-6)).astype(np.float32)
        coef = np.where((det0 <= 1e-6) | (det1 <= 1e-6), np.float32(0), coef)
        opacity = opacity * coef
    d1, d2, off = c00 + ks, c11 + ks, c01
    mid = np.float32(0.5) * (d1 + d2)
    radius = np.sqrt(((d1 - d2) / 2) ** 2 + off ** 2).astype(np.float32)
    l1 = mid + radius
    l2 = np.maximum(mid - radius, np.float32(0.1))
    with np.errstate(all="ignore"):
        dv = np.stack([off, l1 - d1], 1)
        dv = dv / np.linalg.norm(dv, axis=1, keepdims=True).astype(np.float32)
        v1 = np.sqrt(2 * l1)[:, None] * dv
        v2 = np.sqrt(2 * l2)[:, None] * np.stack([dv[:, 1], -dv[:, 0]], 1)
        vc = pos2d[:, :2] / pos2d[:, 3:4]
        cam_pos = view_inv[:3, 3]
        dirv = xyz - cam_pos
        dirv = dirv / np.linalg.norm(dirv, axis=1, keepdims=True).astype(np.float32)
    col = np.maximum(evaluate_sh(dirv.astype(np.float32), sh, int(st.max_sh_deg)), 0)
    vp = np.array(cam.viewport[:], np.float32)
    out = np.concatenate([v1 / vp, v2 / vp, vc, col, opacity[:, None]], 1).astype(np.float32)
    with np.errstate(over="ignore"):
        splats = out.astype(np.float16)
    zfar = -proj[2, 3] / (proj[2, 2] - np.float32(1))
    keys = (zfar - pos2d[:, 2]).astype(np.float32).view(np.uint32)
    return keep, splats, keys
