"""Headless renderer -- what the reference draws into its GGUI window every frame (main.py:51-62, 151-161, 175-188):
camera from the scene file (cam_pos / cam_look_at / cam_up), ambient + one point light, the fluid and the rigid
particles as spheres of radius ps.particle_radius with per-vertex colours.  Here the frame is an RGBA8 image in
device memory (sph_render: depth-tested shaded sphere splats), written as PNG when asked; no window, no display.
"""
import ctypes
import struct
import zlib

import numpy as np
import torch

from . import _lib

# the camera keys the reference requires in every scene file (SURVEY B-15); shipped scenes without them get a view
# of the whole box from the +x +y +z corner
def default_camera(config):
    sc = config["scene"]
    lo, hi = np.array(sc["box_min"], dtype=np.float64), np.array(sc["box_max"], dtype=np.float64)
    centre, size = 0.5 * (lo + hi), float(np.linalg.norm(hi - lo))
    return list(centre + np.array([0.9, 0.6, 1.2]) * size), list(centre), [0.0, 1.0, 0.0]


class Renderer:
    def __init__(self, ps, config, width=640, height=640, background=(26, 26, 26)):
        """640 x 640 is the reference's window resolution (main.py:52)."""
        self.ps, self.width, self.height = ps, int(width), int(height)
        sc = config.get("scene", {})
        pos, look, up = default_camera(config)
        self.cam = _lib.SphCamera()
        for k in range(3):
            self.cam.pos[k] = float(sc.get("cam_pos", pos)[k])            # main.py:60
            self.cam.look_at[k] = float(sc.get("cam_look_at", look)[k])   # main.py:61
            self.cam.up[k] = float(sc.get("cam_up", up)[k])               # main.py:62
            self.cam.light_pos[k] = (0.5, 1.5, 1.5)[k]                    # main.py:154
            self.cam.background[k] = int(background[k])
        self.cam.fov_y_deg = 45.0
        self.cam.ambient = 0.8                                            # main.py:153
        self.render_fluid, self.render_rigid = True, True                 # keys f/g and r/t (main.py:134-141)
        self.image = torch.zeros((self.height, self.width, 4), dtype=torch.uint8, device=ps._device)
        self.depth = torch.zeros((self.height, self.width), dtype=torch.float32, device=ps._device)

    def set_camera(self, pos, look_at, up=(0.0, 1.0, 0.0)):
        for k in range(3):
            self.cam.pos[k], self.cam.look_at[k], self.cam.up[k] = float(pos[k]), float(look_at[k]), float(up[k])

    def frame(self):
        """Draw the current state; returns the device image (H x W x 4 uint8, row 0 = top)."""
        ps = self.ps
        what = (_lib.RENDER_FLUID if self.render_fluid else 0) | (
            _lib.RENDER_RIGID if self.render_rigid and ps.exist_rigid[None] == 1 else 0)
        frgb = ps.fluid_particles.rgb.tensor
        rrgb = ps.rigid_particles.rgb.tensor
        _lib.check(ps._lib.sph_render(ps._h, ctypes.byref(self.cam), self.width, self.height, what,
                                      frgb.data_ptr(), frgb.stride(0), rrgb.data_ptr() if rrgb.numel() else None,
                                      rrgb.stride(0) if rrgb.numel() else 3, self.image.data_ptr(), self.depth.data_ptr(),
                                      ps._stream()), ps._h)
        return self.image

    def save_png(self, path):
        write_png(path, self.frame().cpu().numpy())


def write_png(path, rgba):
    """Minimal PNG writer (8-bit RGBA, zlib from the standard library)."""
    h, w, _ = rgba.shape
    raw = b"".join(b"\x00" + rgba[y].tobytes() for y in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
