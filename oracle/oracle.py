"""ctypes wrapper of the CPU oracle (oracle/ws_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (web-splat_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libws_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "ws_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libws_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


class CameraUniform(C.Structure):       # renderer.rs:290-306, 272 B
    _fields_ = [("view", C.c_float * 16), ("view_inv", C.c_float * 16), ("proj", C.c_float * 16),
                ("proj_inv", C.c_float * 16), ("viewport", C.c_float * 2), ("focal", C.c_float * 2)]


class RenderSettings(C.Structure):      # renderer.rs:604-619, 80 B
    _fields_ = [("clip_min", C.c_float * 4), ("clip_max", C.c_float * 4), ("gaussian_scaling", C.c_float),
                ("max_sh_deg", C.c_uint32), ("mip_splatting", C.c_uint32), ("kernel_size", C.c_float),
                ("walltime", C.c_float), ("scene_extend", C.c_float), ("_pad", C.c_uint32 * 2),
                ("center", C.c_float * 4)]


class Quant(C.Structure):
    _fields_ = [("zero_point", C.c_int32), ("scale", C.c_float), ("_pad", C.c_uint32 * 2)]


class Quant4(C.Structure):
    _fields_ = [("color_dc", Quant), ("color_rest", Quant), ("opacity", Quant), ("scaling_factor", Quant)]


assert C.sizeof(CameraUniform) == 272 and C.sizeof(RenderSettings) == 80 and C.sizeof(Quant4) == 64


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.wso_f32_to_f16.restype = C.c_uint16; L.wso_f32_to_f16.argtypes = [C.c_float]
        L.wso_f16_to_f32.restype = C.c_float; L.wso_f16_to_f32.argtypes = [C.c_uint16]
        L.wso_fit_near_far.restype = None
        L.wso_fit_near_far.argtypes = [vp, vp, vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.wso_camera_uniform_build.restype = None
        L.wso_camera_uniform_build.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float, C.c_float,
                                               C.c_uint32, C.c_uint32, C.POINTER(CameraUniform)]
        L.wso_preprocess_raw.restype = C.c_uint32
        L.wso_preprocess_raw.argtypes = [vp, vp, C.c_uint32, C.POINTER(CameraUniform), C.POINTER(RenderSettings), vp, vp, vp]
        L.wso_preprocess_compressed.restype = C.c_uint32
        L.wso_preprocess_compressed.argtypes = [vp, vp, vp, C.POINTER(Quant4), C.c_uint32, C.c_uint32,
                                                C.POINTER(CameraUniform), C.POINTER(RenderSettings), vp, vp, vp]
        L.wso_sort_pairs.restype = None; L.wso_sort_pairs.argtypes = [vp, vp, C.c_uint32]
        L.wso_composite.restype = None
        L.wso_composite.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp]
        L.wso_tile_rects.restype = None
        L.wso_ply_convert.restype = C.c_int
        L.wso_ply_convert.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp]
        L.wso_c3dgs_convert.restype = C.c_int
        L.wso_c3dgs_convert.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, vp, vp, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32,
                                        C.c_float, C.c_int32, C.c_float, C.c_int32, vp, vp, vp, vp, vp, vp]
        L.wso_tile_rects.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.POINTER(C.c_uint64)]
        L.wso_num_threads.restype = C.c_int
        L.wso_set_num_threads.restype = None; L.wso_set_num_threads.argtypes = [C.c_int]
        L.wso_composite_rop.restype = None
        L.wso_composite_rop.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_int, vp]
        _lib = L
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def f32_to_f16_bits(x):
    return lib().wso_f32_to_f16(float(np.float32(x)))


def f16_bits_to_f32(h):
    return lib().wso_f16_to_f32(int(h))


def num_threads():
    return lib().wso_num_threads()


def set_num_threads(n):
    """OpenMP threads of the oracle (torchrun exports OMP_NUM_THREADS=1: bench.py sets the count itself)."""
    lib().wso_set_num_threads(int(n))


def host_cores():
    """Cores this process may run on (cgroup / affinity aware)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def aabb_radius(bmin, bmax):
    d = np.asarray(bmax, np.float32) - np.asarray(bmin, np.float32)
    r2 = np.float32(0)
    for t in d:
        r2 = np.float32(r2 + np.float32(t * t))
    return np.float32(np.sqrt(r2) / np.float32(2.0))


def fit_near_far(pos, bmin, bmax):
    pos = np.ascontiguousarray(pos, np.float32); bmin = np.ascontiguousarray(bmin, np.float32); bmax = np.ascontiguousarray(bmax, np.float32)
    zn, zf = C.c_float(), C.c_float()
    lib().wso_fit_near_far(_p(pos), _p(bmin), _p(bmax), C.byref(zn), C.byref(zf))
    return zn.value, zf.value


def camera_uniform(pos, rot_wxyz, fovx, fovy, znear, zfar, W, H):
    pos = np.ascontiguousarray(pos, np.float32); rot = np.ascontiguousarray(rot_wxyz, np.float32)
    u = CameraUniform()
    lib().wso_camera_uniform_build(_p(pos), _p(rot), fovx, fovy, znear, zfar, W, H, C.byref(u))
    return u


def render_settings(cloud, gaussian_scaling=1.0, max_sh_deg=3, mip_splatting=None, kernel_size=None,
                    clipping_box=None, walltime=100.0, scene_extend=None,
                    pc_mip=None, pc_kernel=None):
    """SplattingArgsUniform::from_args_and_pc, renderer.rs:620-651."""
    s = RenderSettings()
    s.gaussian_scaling = gaussian_scaling
    s.max_sh_deg = max_sh_deg
    s.mip_splatting = int(mip_splatting if mip_splatting is not None else bool(pc_mip))
    s.kernel_size = kernel_size if kernel_size is not None else (pc_kernel if pc_kernel is not None else 0.3)
    if clipping_box is None:
        lo, hi = cloud["aabb_min"], cloud["aabb_max"]
    elif hasattr(clipping_box, "min"):                      # an Aabb-like object
        lo, hi = clipping_box.min, clipping_box.max
    else:
        lo, hi = clipping_box
    for i in range(3):
        s.clip_min[i] = float(lo[i]); s.clip_max[i] = float(hi[i]); s.center[i] = float(cloud["center"][i])
    s.walltime = walltime
    rad = float(aabb_radius(cloud["aabb_min"], cloud["aabb_max"]))
    ext = rad if scene_extend is None else float(np.float32(scene_extend))
    s.scene_extend = ext if ext > rad else rad
    return s


def preprocess(cloud, cam, settings, file_sh_deg=None):
    """Stage 1.  Returns (splats [V,10] u16, keys [V] u32, src [V] u32 Gaussian index per slot)."""
    n = int(cloud["num_points"])
    splats = np.zeros((max(n, 1), 10), np.uint16); keys = np.zeros(max(n, 1), np.uint32); src = np.zeros(max(n, 1), np.uint32)
    g = np.ascontiguousarray(cloud["gaussians"]).view(np.uint8).reshape(-1)
    sh = np.ascontiguousarray(cloud["sh_coefs"]).view(np.uint8).reshape(-1)
    if not cloud["compressed"]:
        v = lib().wso_preprocess_raw(_p(g), _p(sh), n, C.byref(cam), C.byref(settings), _p(splats), _p(keys), _p(src))
    else:
        cov = np.ascontiguousarray(cloud["covars"]).view(np.uint8).reshape(-1)
        q = Quant4()
        for name in ("color_dc", "color_rest", "opacity", "scaling_factor"):
            zp, sc = cloud["quantization"][name]
            getattr(q, name).zero_point = int(zp); getattr(q, name).scale = float(sc)
        deg = cloud["sh_deg"] if file_sh_deg is None else file_sh_deg
        v = lib().wso_preprocess_compressed(_p(g), _p(sh), _p(cov), C.byref(q), n, deg, C.byref(cam), C.byref(settings),
                                            _p(splats), _p(keys), _p(src))
    return splats[:v].copy(), keys[:v].copy(), src[:v].copy()


def sort_pairs(keys, vals):
    """Stage 2: stable ascending (u32 key, u32 payload) sort, returns new arrays."""
    k = np.ascontiguousarray(keys, np.uint32).copy(); v = np.ascontiguousarray(vals, np.uint32).copy()
    lib().wso_sort_pairs(_p(k), _p(v), k.size)
    return k, v


def composite(splats, order, W, H, clear=(0, 0, 0, 0), want_sens=False):
    """Stage 3: back-to-front 'over' of the splats in `order`.  Returns f32 [H,W,4] (+ sens [H,W])."""
    splats = np.ascontiguousarray(splats, np.uint16); order = np.ascontiguousarray(order, np.uint32)
    out = np.empty((H, W, 4), np.float32)
    sens = np.empty((H, W), np.float32) if want_sens else None
    clr = np.ascontiguousarray(clear, np.float32)
    lib().wso_composite(_p(splats), _p(order), order.size, W, H, _p(clr), _p(out), _p(sens) if want_sens else None)
    return (out, sens) if want_sens else out


def composite_rop(splats, order, W, H, fmt, clear=(0, 0, 0, 0)):
    """Stage 3 as the reference's render target would hold it: the blend result is rounded to the target format
    (fmt 0 = Rgba8Unorm, 1 = Rgba16Float, 2 = Rgba32Float) after EVERY layer (renderer.rs:63-67).  f32 [H,W,4]."""
    splats = np.ascontiguousarray(splats, np.uint16); order = np.ascontiguousarray(order, np.uint32)
    out = np.empty((H, W, 4), np.float32)
    clr = np.ascontiguousarray(clear, np.float32)
    lib().wso_composite_rop(_p(splats), _p(order), order.size, W, H, _p(clr), int(fmt), _p(out))
    return out


def tile_rects(splats, W, H):
    """New-design intermediate: inclusive tile rect {x0,y0,x1,y1} per stored splat and the pair count P."""
    splats = np.ascontiguousarray(splats, np.uint16)
    r = np.zeros((max(len(splats), 1), 4), np.int32)
    tot = C.c_uint64()
    lib().wso_tile_rects(_p(splats), len(splats), W, H, _p(r), C.byref(tot))
    return r[:len(splats)], tot.value


def render_frame(cloud, pos, rot_wxyz, W, H, fovx, fovy, clear=(0, 0, 0, 0), want_sens=False, **settings_kw):
    """Whole reference frame as bin/render.rs:55-127 does it: fit_near_far, prepare, render."""
    zn, zf = fit_near_far(pos, cloud["aabb_min"], cloud["aabb_max"])
    cam = camera_uniform(pos, rot_wxyz, fovx, fovy, zn, zf, W, H)
    st = render_settings(cloud, **settings_kw)
    splats, keys, _ = preprocess(cloud, cam, st)
    _, order = sort_pairs(keys, np.arange(len(keys), dtype=np.uint32))
    img = composite(splats, order, W, H, clear, want_sens)
    return dict(image=img[0] if want_sens else img, sens=img[1] if want_sens else None,
                splats=splats, keys=keys, order=order, cam=cam, settings=st)


def ply_convert(vertices, sh_deg):
    """vertices: (n, floats_per_vertex) float32 in the 3DGS .ply property order.
    Returns dict(gaussians (n,28) u8, sh_coefs (n,96) u8, bbox[6], center[3], up or None)."""
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    n, stride = v.shape
    assert stride == 14 + 3 * (sh_deg + 1) ** 2
    g = np.zeros((n, 28), np.uint8)
    sh = np.zeros((n, 96), np.uint8)
    bbox = np.zeros(6, np.float32); center = np.zeros(3, np.float32); up = np.zeros(3, np.float32)
    with np.errstate(all="ignore"):
        has_up = lib().wso_ply_convert(_p(v), n, stride, sh_deg, _p(g), _p(sh), _p(bbox), _p(center), _p(up))
    return dict(gaussians=g, sh_coefs=sh, bbox=bbox, center=center, up=up if has_up else None)


def c3dgs_convert(arrays):
    """arrays: dict of the .npz members (io/npz.rs:58-160).  Returns gaussians (n,24) u8, sh_coefs i8 (K, 3C),
    covars (Kc, 12) u8, bbox[6], center[3], up or None."""
    def opt(name, dt):
        return np.ascontiguousarray(arrays[name], dtype=dt) if arrays.get(name) is not None else None
    xyz = np.ascontiguousarray(arrays["xyz"], dtype=np.float16).reshape(-1, 3)
    n = len(xyz)
    opacity = np.ascontiguousarray(arrays["opacity"], np.int8).reshape(-1)
    sf, gi, fi = opt("scaling_factor", np.int8), opt("gaussian_indices", np.int32), opt("feature_indices", np.int32)
    scaling = np.ascontiguousarray(arrays["scaling"], np.int8).reshape(-1, 3)
    rotation = np.ascontiguousarray(arrays["rotation"], np.int8).reshape(-1, 4)
    dc = np.ascontiguousarray(arrays["features_dc"], np.int8).reshape(-1, 3)
    rest = np.ascontiguousarray(arrays["features_rest"], np.int8)
    sh_deg = int(round((rest.shape[1] + 1) ** 0.5)) - 1 if rest.ndim == 3 else 0
    per = 3 * (sh_deg + 1) ** 2
    g = np.zeros((n, 24), np.uint8); sh = np.zeros((len(dc), per), np.int8); cov = np.zeros((len(scaling), 12), np.uint8)
    bbox = np.zeros(6, np.float32); center = np.zeros(3, np.float32); up = np.zeros(3, np.float32)
    pn = lambda a: _p(a) if a is not None else None
    with np.errstate(all="ignore"):
        has_up = lib().wso_c3dgs_convert(_p(xyz.view(np.uint16)), _p(opacity), pn(sf), pn(gi), pn(fi), n, _p(scaling), _p(rotation), len(scaling),
                                         _p(dc), _p(rest), len(dc), sh_deg,
                                         float(arrays.get("scaling_scale", 1.0)), int(arrays.get("scaling_zero_point", 0)),
                                         float(arrays.get("rotation_scale", 1.0)), int(arrays.get("rotation_zero_point", 0)),
                                         _p(g), _p(sh), _p(cov), _p(bbox), _p(center), _p(up))
    return dict(gaussians=g, sh_coefs=sh, covars=cov, sh_deg=sh_deg, bbox=bbox, center=center, up=up if has_up else None)
