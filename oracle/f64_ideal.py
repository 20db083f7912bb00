"""Float64 "ideal" splat renderer: an INDEPENDENT cross-check of the f16-faithful oracle.

TEST INFRASTRUCTURE ONLY (see oracle/ws_oracle.c header).  ws_oracle.c restates the reference's shaders line by
line (f32 arithmetic, f16-packed 2D splats, quad + `a = p.p` fragment test).  This file does NOT follow the
shaders' structure: it renders the same scene from the textbook 3D-Gaussian-splatting formulation in float64,
with no f16 intermediate and no eigen-axis quad:

    pinhole camera, +z forward, +y down:  u = fx x/z + W/2,  v = fy y/z + H/2          (camera.rs:216-242)
    J = d(u,v)/d(x,y,z),  S2 = J R S3 R^T J^T + k I  (pixels^2)                        (preprocess.wgsl:209-223,238)
    eigenvalues of S2 with the smaller one floored at 0.1                               (preprocess.wgsl:243-245)
    a(d) = 1/2 d^T S2'^-1 d,  covered iff a <= 2*CUTOFF,  alpha = min(0.99, o exp(-a))  (gaussian.wgsl:59-66)
    colour = max(0, SH(dir) + 0.5)                                                      (preprocess.wgsl:255-260)
    'over' compositing far -> near by camera depth                                      (renderer.rs:63-67)

The identities a = p.p = 1/2 d^T S2'^-1 d and "NDC y up + flipped projection = pixel y down" are derived in
SURVEY.md appendix A.3; if the C restatement (or the CUDA path) misread a transpose, a sign or the cutoff, the two
renderers disagree far beyond the f16-rounding differences that legitimately separate them.
"""
import math

import numpy as np

CUTOFF = 2.3539888583335364          # gaussian.wgsl:2
_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


def _quat_to_world2cam(q):
    """cgmath Matrix3::from(Quaternion(w, x, y, z)) as a math matrix (row, col): the world->camera rotation."""
    w, x, y, z = [float(t) for t in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], np.float64)


def _sh_colour(d, sh, deg):
    x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    c = _C0 * sh[:, 0]
    if deg > 0:
        c = c - _C1 * y * sh[:, 1] + _C1 * z * sh[:, 2] - _C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        c = c + _C2[0] * xy * sh[:, 4] + _C2[1] * yz * sh[:, 5] + _C2[2] * (2 * zz - xx - yy) * sh[:, 6] \
              + _C2[3] * xz * sh[:, 7] + _C2[4] * (xx - yy) * sh[:, 8]
    if deg > 2:
        c = c + _C3[0] * y * (3 * xx - yy) * sh[:, 9] + _C3[1] * xy * z * sh[:, 10] + _C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] \
              + _C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12] + _C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] \
              + _C3[5] * z * (xx - yy) * sh[:, 14] + _C3[6] * x * (xx - 3 * yy) * sh[:, 15]
    return np.maximum(c + 0.5, 0.0)


def render(cloud, pos, rot_wxyz, W, H, fovx, fovy, znear, zfar, clear=(0.0, 0.0, 0.0, 0.0), kernel_size=0.3, max_sh_deg=3):
    """Raw-layout cloud dict (web-splat_b200.synth) -> float64 image [H, W, 4], plus the per-Gaussian projected centre."""
    g = cloud["gaussians"]
    mu = g["xyz"].astype(np.float64)
    n = len(mu)
    cov6 = g["cov"].astype(np.float64)
    opac = g["opacity"].astype(np.float64)
    sh = cloud["sh_coefs"].astype(np.float64)
    S3 = np.empty((n, 3, 3))
    S3[:, 0, 0] = cov6[:, 0]; S3[:, 0, 1] = S3[:, 1, 0] = cov6[:, 1]; S3[:, 0, 2] = S3[:, 2, 0] = cov6[:, 2]
    S3[:, 1, 1] = cov6[:, 3]; S3[:, 1, 2] = S3[:, 2, 1] = cov6[:, 4]; S3[:, 2, 2] = cov6[:, 5]
    R = _quat_to_world2cam(rot_wxyz)
    eye = np.asarray(pos, np.float64)
    pc = (mu - eye) @ R.T                                   # camera space
    fx = W / (2 * math.tan(fovx / 2)); fy = H / (2 * math.tan(fovy / 2))
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    lo, hi = np.asarray(cloud["aabb_min"], np.float64), np.asarray(cloud["aabb_max"], np.float64)
    with np.errstate(all="ignore"):
        keep = ((mu >= lo) & (mu <= hi)).all(1) & (z > znear) & (z < zfar)             # clip box, 0 < z_ndc < 1
        keep &= (np.abs(x / z) <= 1.2 * math.tan(fovx / 2)) & (np.abs(y / z) <= 1.2 * math.tan(fovy / 2))
    img = np.empty((H, W, 4)); img[:] = np.asarray(clear, np.float64)
    idx = np.nonzero(keep)[0]
    if len(idx) == 0:
        return img, idx
    x, y, z, pc = x[idx], y[idx], z[idx], pc[idx]
    J = np.zeros((len(idx), 2, 3))
    J[:, 0, 0] = fx / z; J[:, 0, 2] = -fx * x / (z * z)
    J[:, 1, 1] = fy / z; J[:, 1, 2] = -fy * y / (z * z)
    M = J @ R
    S2 = M @ S3[idx] @ np.transpose(M, (0, 2, 1))
    S2[:, 0, 0] += kernel_size; S2[:, 1, 1] += kernel_size
    lam, vec = np.linalg.eigh(S2)                           # ascending eigenvalues
    lam[:, 0] = np.maximum(lam[:, 0], 0.1)                  # lambda2 floor
    Sinv = vec @ (np.eye(2)[None] / lam[:, None, :]) @ np.transpose(vec, (0, 2, 1))
    u = fx * x / z + W / 2; v = fy * y / z + H / 2
    dirs = mu[idx] - eye
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    col = _sh_colour(dirs, sh[idx], max_sh_deg)
    o = opac[idx]
    # far -> near; equal depths in index order (the oracle's deterministic tie-break)
    order = np.lexsort((idx, -z))
    thr = 2 * CUTOFF
    rad = np.sqrt(2 * thr * lam[:, 1])                      # |d| <= sqrt(2 thr lambda_max) is necessary for a <= thr
    for i in order:
        x0 = max(int(math.floor(u[i] - rad[i] - 1)), 0); x1 = min(int(math.ceil(u[i] + rad[i] + 1)), W - 1)
        y0 = max(int(math.floor(v[i] - rad[i] - 1)), 0); y1 = min(int(math.ceil(v[i] + rad[i] + 1)), H - 1)
        if x1 < x0 or y1 < y0:
            continue
        dx = (np.arange(x0, x1 + 1) + 0.5 - u[i])[None, :]
        dy = (np.arange(y0, y1 + 1) + 0.5 - v[i])[:, None]
        a = 0.5 * (Sinv[i, 0, 0] * dx * dx + 2 * Sinv[i, 0, 1] * dx * dy + Sinv[i, 1, 1] * dy * dy)
        b = np.where(a <= thr, np.minimum(0.99, o[i] * np.exp(-a)), 0.0)
        blk = img[y0:y1 + 1, x0:x1 + 1]
        blk[..., :3] = col[i][None, None, :] * b[..., None] + blk[..., :3] * (1 - b[..., None])
        blk[..., 3] = b + blk[..., 3] * (1 - b)
    return img, idx
