"""cameras.json, camera conversion, pixel conversion and the PNG writer (host logic),
and the offline dataset renderer end to end on the GPU."""
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest


def _scene_json(ws, count=17, W=320, H=200):
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    cams = []
    for i in range(count):
        pos, rot = ws.synth.orbit_camera(360.0 * i / count, radius=3.0 + 0.1 * i)
        cam = ws.PerspectiveCamera(pos, rot, ws.PerspectiveProjection(fovx, fovy, 0.01, 100.0))
        cams.append(ws.SceneCamera.from_perspective(ws, cam, "img_%03d" % i, i, (W, H)).to_json())
    return cams


def test_scene_from_json_split_and_queries(ws):
    entries = _scene_json(ws)
    sc = ws.Scene.from_json(json.dumps(entries))
    assert sc.num_cameras() == 17
    test, train = sc.cameras(ws.scene.TEST), sc.cameras(ws.scene.TRAIN)
    assert [c.id for c in test] == [0, 8, 16] and len(train) == 14          # scene.rs:140-147: every 8th view is test
    assert [c.id for c in sc.cameras()] == list(range(17))
    assert sc.camera(3).img_name == "img_003" and sc.camera(99) is None
    pos = np.array([e["position"] for e in entries], np.float64)
    want = max(np.linalg.norm(a - b) for a in pos for b in pos)
    assert math.isclose(sc.extend(), want, rel_tol=1e-5)
    assert sc.nearest_camera(entries[5]["position"]) == 5
    assert sc.nearest_camera(entries[5]["position"], ws.scene.TEST) in (0, 8)
    # duplicate ids: the later entry wins, like HashMap::insert
    dup = entries + [dict(entries[2], img_name="again")]
    assert ws.Scene.from_json(json.dumps(dup)).camera(2).img_name == "again"


def test_scene_camera_round_trip(ws):
    """from_perspective -> Into<PerspectiveCamera> returns the camera (scene.rs:40-108)."""
    W, H = 640, 360
    fovx, fovy = ws.synth.fov_for_viewport(W, H)
    for az in (0.0, 77.0, 181.0, 290.0):
        pos, rot = ws.synth.orbit_camera(az)
        cam = ws.PerspectiveCamera(pos, rot, ws.PerspectiveProjection(fovx, fovy, 0.01, 100.0))
        back = ws.SceneCamera.from_perspective(ws, cam, "x", 0, (W, H)).to_perspective(ws)
        assert np.allclose(back.position, pos)
        q = back.rotation if np.dot(back.rotation, rot) > 0 else -back.rotation
        assert np.allclose(q, rot, atol=2e-6)
        assert math.isclose(back.projection.fovx, fovx, rel_tol=1e-6) and math.isclose(back.projection.fovy, fovy, rel_tol=1e-6)
        assert back.projection.znear == 0.01 and back.projection.zfar == 100.0
        assert math.isclose(back.projection.fov2view_ratio, (W / H) / (fovx / fovy), rel_tol=1e-6)
    # a left-handed rotation in the file gets its y row flipped (scene.rs:89-95)
    sc = ws.SceneCamera(0, "x", W, H, [0, 0, 0], np.diag([1.0, -1.0, 1.0]), 500.0, 500.0)
    assert np.allclose(np.abs(sc.to_perspective(ws).rotation), [1, 0, 0, 0])
    assert math.isclose(ws.scene.focal2fov(500.0, 1000.0), 2 * math.atan(1.0), rel_tol=1e-6)
    assert math.isclose(ws.scene.fov2focal(ws.scene.focal2fov(432.1, 800.0), 800.0), 432.1, rel_tol=1e-5)


def test_pixel_conversion_and_png(ws):
    f = np.array([[[-1.0, 0.0, 0.5, 1.0], [2.0, 0.999, 1 / 255, np.nan]]], np.float16)
    px = ws.scene.frame_to_rgba8(f)
    assert px.tolist() == [[[0, 0, 127, 255], [255, 254, 0, 0]]]              # truncation, not rounding (bin/render.rs:237)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(37, 53, 4), dtype=np.uint8)
    data = ws.scene.png_bytes(img)
    assert np.array_equal(ws.scene.decode_png(data), img)
    # f32: 1920/1600 = 1.2000000477, 1080/that = 899.99996 -> 899 (the reference's own arithmetic, bin/render.rs:59-62)
    assert ws.scene.render_resolution(1920, 1080) == (1600, 899) and ws.scene.render_resolution(1600, 901) == (1600, 901)
    assert ws.scene.render_resolution(4946, 3286) == (1600, int(np.float32(3286) / (np.float32(4946) / np.float32(1600))))


@pytest.mark.gpu
def test_gpu_offline_renderer_end_to_end(ws, orc, ctx, tmp_path):
    """.ply + cameras.json -> PNGs through scripts/render_scene.py; one view re-rendered through the API and the oracle."""
    n, W, H = 20000, 320, 200
    v = ws.synth.ply_vertices(n, 8, 3)
    ply = tmp_path / "cloud.ply"; ply.write_bytes(ws.synth.ply_bytes(v, 3))
    entries = _scene_json(ws, 9, W, H)
    cams = tmp_path / "cameras.json"; cams.write_text(json.dumps(entries))
    out = tmp_path / "out"
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    p = subprocess.run([sys.executable, os.path.join(root, "scripts", "render_scene.py"), str(ply), str(cams), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert sorted(os.listdir(out / "test")) == ["00000.png", "00001.png"]      # views 0 and 8
    assert len(os.listdir(out / "train")) == 7
    # re-render train view 2 (= camera id 3) and compare with its PNG and with the oracle
    sc = ws.Scene.from_json(str(cams))
    s = sc.cameras(ws.scene.TRAIN)[2]
    assert s.id == 3
    pc = ws.PointCloud.from_ply(ctx, ply.read_bytes())
    cam = s.to_perspective(ws); cam.fit_near_far(pc.bbox())
    r = ws.GaussianRenderer.new(ctx, ws.FORMAT_RGBA16_FLOAT, 3, False)
    r.prepare(None, pc, ws.SplattingArgs(cam, (W, H), max_sh_deg=3))
    frame = r.empty_host_frame(); r.render_to_host(frame, pc)
    import torch
    torch.cuda.synchronize()
    png = ws.scene.decode_png((out / "train" / "00002.png").read_bytes())
    assert np.array_equal(png, ws.scene.frame_to_rgba8(frame)) and png.any()
    o = orc.ply_convert(v, 3)
    cloud = dict(gaussians=o["gaussians"].view(ws.synth.GAUSSIAN_DTYPE).reshape(-1), sh_coefs=o["sh_coefs"].view(np.float16).reshape(-1, 16, 3),
                 num_points=n, sh_deg=3, compressed=False, aabb_min=o["bbox"][:3], aabb_max=o["bbox"][3:], center=o["center"])
    ref = orc.render_frame(cloud, cam.position, cam.rotation, W, H, cam.projection.fovx, cam.projection.fovy)
    # the frame is f16 (half an ulp at 0.5 is 1.2e-4), the oracle image f32
    assert np.abs(frame.astype(np.float32) - ref["image"]).mean() < 1.5e-4
    assert np.abs(png.astype(np.int32) - ws.scene.frame_to_rgba8(ref["image"].astype(np.float16)).astype(np.int32)).max() <= 2
