"""Headless renderer (main.py:151-161): projection, sphere footprint, depth test, colours."""
import contextlib
import io
import json
import struct
import zlib

import numpy as np
import pytest

from cfd_taichi_b200 import scenes
from cfd_taichi_b200.render import Renderer, write_png
from conftest import quiet_ps

pytestmark = pytest.mark.gpu


def test_two_particles_projection_and_depth(built):
    import torch
    cfg = scenes.shipped("small_block", "wcsph")
    ps = quiet_ps(cfg, solver_name="wcsph")
    n = ps.particle_num
    # park everything far behind the camera, then place two particles on the optical axis
    ps._pos4[:n, :3] = torch.tensor([0.0, 0.0, 100.0], device=ps._device)
    ps._pos4[0, :3] = torch.tensor([0.0, 0.0, -2.0], device=ps._device)   # near: red
    ps._pos4[1, :3] = torch.tensor([0.0, 0.0, -4.0], device=ps._device)   # far, same pixel: green, must be hidden
    ps._pos4[2, :3] = torch.tensor([0.5, 0.0, -2.0], device=ps._device)   # off axis: blue
    rgb = ps.fluid_particles.rgb.tensor
    rgb[0] = torch.tensor([1.0, 0.0, 0.0], device=ps._device)
    rgb[1] = torch.tensor([0.0, 1.0, 0.0], device=ps._device)
    rgb[2] = torch.tensor([0.0, 0.0, 1.0], device=ps._device)
    r = Renderer(ps, cfg, width=400, height=400, background=(0, 0, 0))
    r.set_camera([0, 0, 0], [0, 0, -1], [0, 1, 0])
    r.cam.light_pos[0], r.cam.light_pos[1], r.cam.light_pos[2] = 0.0, 0.0, 0.0   # headlight: full shading at the centre
    img = r.frame().cpu().numpy()
    depth = r.depth.cpu().numpy()
    focal = 200.0 / np.tan(np.radians(22.5))
    # centre pixel: the near (red) particle, not the far green one; depth = distance to the sphere's front
    c = img[200, 200]
    assert c[0] > 200 and c[1] == 0 and c[2] == 0 and c[3] == 255
    assert abs(depth[200, 200] - (2.0 - 0.025)) < 2e-3
    # its footprint is a disc of radius focal * r / z pixels
    red = (img[:, :, 0] > 0) & (img[:, :, 1] == 0) & (img[:, :, 2] == 0)
    rad = focal * 0.025 / 2.0
    assert abs(red.sum() - np.pi * rad * rad) < 0.2 * np.pi * rad * rad
    ys, xs = np.nonzero(red)
    assert abs(xs.mean() - 199.5) < 1.0 and abs(ys.mean() - 199.5) < 1.0
    # nothing green anywhere (fully occluded); the blue one sits focal * 0.5 / 2 pixels to the right
    assert not ((img[:, :, 1] > 0) & (img[:, :, 0] == 0)).any()
    blue = (img[:, :, 2] > 0) & (img[:, :, 0] == 0)
    ys, xs = np.nonzero(blue)
    assert abs(xs.mean() - (199.5 + focal * 0.5 / 2.0)) < 1.5 and abs(ys.mean() - 199.5) < 1.0
    # background elsewhere, depth +inf
    assert (img[0, 0] == np.array([0, 0, 0, 255])).all() and np.isinf(depth[0, 0])
    ps.close()


def test_scene_frame_and_png(built, tmp_path):
    cfg = scenes.shipped("small_block", "wcsph")
    cfg["scene"].update(cam_pos=[3.2, 2.2, 3.6], cam_look_at=[0.65, 1.1, 0.65], cam_up=[0, 1, 0])
    ps = quiet_ps(cfg, solver_name="wcsph")
    r = Renderer(ps, cfg)
    img = r.frame().cpu().numpy()
    assert img.shape == (640, 640, 4)
    drawn = (img[:, :, :3] != np.array([26, 26, 26])).any(axis=2)
    assert 2000 < drawn.sum() < 640 * 640 // 2          # the block is in view and does not fill the frame
    assert img[drawn][:, 2].mean() > img[drawn][:, 0].mean() + 50   # the fluid's blue (0, 0.28, 1) dominates
    path = tmp_path / "frame.png"
    r.save_png(str(path))
    data = path.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    w, h = struct.unpack(">II", data[16:24])
    assert (w, h) == (640, 640)
    # decode the IDAT stream back and compare with the frame
    pos, idat = 8, b""
    while pos < len(data):
        ln, tag = struct.unpack(">I", data[pos:pos + 4])[0], data[pos + 4:pos + 8]
        if tag == b"IDAT":
            idat += data[pos + 8:pos + 8 + ln]
        pos += 12 + ln
    raw = zlib.decompress(idat)
    rows = np.frombuffer(raw, dtype=np.uint8).reshape(640, 1 + 640 * 4)
    assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(640, 640, 4), img)
    ps.close()


def test_main_writes_rendered_frames(built, tmp_path):
    from cfd_taichi_b200 import main as app
    cfg = scenes.shipped("small_block", "wcsph")
    cfg["scene"]["output_fps"] = 1000
    p = tmp_path / "scene.json"
    p.write_text(json.dumps(cfg))
    with contextlib.redirect_stdout(io.StringIO()):
        ps, solver, rs, t = app.run(app.utils.read_config(str(p)), max_frames=4, quiet=True, render_dir=str(tmp_path / "frames"))
    assert len(sorted((tmp_path / "frames").glob("frame_*.png"))) >= 2
    ps.close()
