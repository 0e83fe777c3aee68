"""Writes the scene files a user of the reference expects to find: default.json, config/<name>.json for every
scene of cfd_taichi_b200/scenes.py (file names as in the reference's config/ directory), BASELINE.json's synthetic
blocks, and the two box meshes obj/cube1.STL (0.8 x 0.5 x 1.0) and obj/cube2.STL (unit cube) as binary STL.

    python config/make_configs.py
"""
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cfd_taichi_b200 import scenes  # noqa: E402


def write_box_stl(path, size):
    """Binary STL of the axis-aligned box [0, size]: 12 outward-facing triangles."""
    sx, sy, sz = (np.float32(v) for v in size)
    v = np.array([[x, y, z] for x in (0, sx) for y in (0, sy) for z in (0, sz)], dtype=np.float32)
    quads = [((0, 1, 3, 2), (-1, 0, 0)), ((4, 6, 7, 5), (1, 0, 0)), ((0, 4, 5, 1), (0, -1, 0)),
             ((2, 3, 7, 6), (0, 1, 0)), ((0, 2, 6, 4), (0, 0, -1)), ((1, 5, 7, 3), (0, 0, 1))]
    with open(path, "wb") as f:
        f.write(b"box mesh written by config/make_configs.py".ljust(80, b" "))
        f.write(struct.pack("<I", 12))
        for (a, b, c, d), n in quads:
            for tri in ((a, b, c), (a, c, d)):
                f.write(struct.pack("<3f", *n))
                for k in tri:
                    f.write(struct.pack("<3f", *v[k]))
                f.write(struct.pack("<H", 0))


def main():
    os.makedirs(os.path.join(ROOT, "obj"), exist_ok=True)
    write_box_stl(os.path.join(ROOT, "obj", "cube1.STL"), (0.8, 0.5, 1.0))
    write_box_stl(os.path.join(ROOT, "obj", "cube2.STL"), (1.0, 1.0, 1.0))
    for fname, (scene, solver) in scenes.FILES.items():
        cfg = scenes.shipped(scene, solver)
        path = os.path.join(ROOT, "default.json" if fname == "default" else os.path.join("config", fname + ".json"))
        with open(path, "w") as f:
            json.dump(cfg, f, indent=1, sort_keys=True)
            f.write("\n")
    # BASELINE.json configs 1-3 as files (config 0 / 3 are breaking_dam_30k / dam_flush_cube with solver.name overridden)
    for fname, cfg in (("dam_30k_wcsph", scenes.shipped("breaking_dam_30k", "wcsph")),
                       ("dam_flush_cube_dfsph", scenes.shipped("dam_flush_cube", "dfsph")),
                       ("dam_1m_dfsph", scenes.breaking_dam(100)),
                       ("dam_4m_pcisph", scenes.breaking_dam(160, "pcisph", 1.5e-4)),
                       ("dam_4m_iisph", scenes.breaking_dam(160, "iisph", 2.5e-4)),
                       ("dam_8m_dfsph", scenes.breaking_dam(200))):
        with open(os.path.join(ROOT, "config", fname + ".json"), "w") as f:
            json.dump(cfg, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()
