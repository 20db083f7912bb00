"""How far is a float compositor (f32 accumulation, ONE conversion to the target format at the end -- what this
repo's CUDA path does, and what wso_composite + conversion does) from the reference's own render target, which holds
the blend result ROUNDED TO THE TARGET FORMAT AFTER EVERY LAYER (renderer.rs:63-67; Rgba8Unorm lib.rs:192-196 /
measure.rs:184, Rgba16Float render.rs:154)?  CPU only (oracle); writes profiles/r02_rop_gap.json.
usage: python tests/measure_rop_gap.py [cfg1 cfg2 ...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import websplat_b200 as ws                 # noqa: E402
from oracle import oracle as orc           # noqa: E402


def to_format(img, fmt):
    """the float compositor's single final conversion (composite.cu epilogue)"""
    if fmt == 0:
        return np.rint(np.clip(img, 0.0, 1.0) * 255.0).astype(np.float32) / np.float32(255.0)
    if fmt == 1:
        return img.astype(np.float16).astype(np.float32)
    return img


def stats(d):
    return {"max": float(d.max()), "mean": float(d.mean()), "p99": float(np.quantile(d, 0.99)), "p999": float(np.quantile(d, 0.999))}


def main():
    out = {}
    for cfg in (sys.argv[1:] or ["cfg1", "cfg2"]):
        n, W, H, seed, _ = ws.synth.CONFIGS[cfg]
        cloud = ws.synth.make_cloud(n, seed)
        pos, rot = ws.synth.fixed_camera() if cfg == "cfg1" else ws.synth.orbit_camera(40.0)
        fovx, fovy = ws.synth.fov_for_viewport(W, H)
        fr = orc.render_frame(cloud, pos, rot, W, H, fovx, fovy)
        res = {"N": n, "W": W, "H": H, "V": int(len(fr["keys"]))}
        hdr = (fr["image"][..., :3] > 1.0).any(axis=2)           # pixels whose FINAL colour exceeds the unorm range
        for fmt, name in ((0, "rgba8unorm"), (1, "rgba16float")):
            rop = orc.composite_rop(fr["splats"], fr["order"], W, H, fmt)
            flt = to_format(fr["image"], fmt)
            d = np.abs(flt - rop).max(axis=2)
            r = {"all_pixels": stats(d)}
            if fmt == 0:
                r["unit"] = "1 = full scale; 1/255 = 0.00392"
                r["pixels_final_colour_le_1"] = stats(d[~hdr]) if (~hdr).any() else None
                r["fraction_pixels_final_colour_gt_1"] = float(hdr.mean())
                r["fraction_gt_2_steps"] = float((d > 2.0 / 255.0 + 1e-7).mean())
                r["fraction_gt_4_steps"] = float((d > 4.0 / 255.0 + 1e-7).mean())
                # LDR variant: the same scene with its colours clamped to [0, 1] in the 2D splats (what a trained scene mostly is)
                sp = fr["splats"].copy()
                c = sp[:, 6:9].view(np.float16)
                c[:] = np.minimum(c, np.float16(1.0))
                img_ldr = orc.composite(sp, fr["order"], W, H)
                rop_ldr = orc.composite_rop(sp, fr["order"], W, H, 0)
                r["ldr_scene_colours_clamped_to_1"] = stats(np.abs(to_format(img_ldr, 0) - rop_ldr).max(axis=2))
            else:
                rel = d / np.maximum(np.abs(rop).max(axis=2), 2.0 ** -10)
                r["relative_to_pixel_max"] = stats(rel)
            res[name] = r
        out[cfg] = res
        print(cfg, json.dumps(res), flush=True)
    with open(os.path.join(ROOT, "profiles", "r02_rop_gap.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
