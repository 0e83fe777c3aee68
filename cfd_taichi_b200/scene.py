"""Host-side scene logic of ParticleSystem.__init__ (reference ParticleSystem.py:31-223): derived
sizes, lattice / boundary-shell initialisation and mesh voxelisation.  Pure numpy, no GPU: this is
the part of the drop-in boundary that runs once before the hot path starts.

Arithmetic follows the reference literally: Python-scope expressions are evaluated in fp64 and cast
to f32 where they meet an f32 Taichi expression; kernel arithmetic is done in np.float32.
"""
import math
import os
import struct

import numpy as np

F32 = np.float32


def derive_sizes(config):
    """particle_num (PS:85-86), boundary_particles_num (PS:129-137), grid_num (PS:100-101)."""
    scene, fluid = config["scene"], config["fluid"]
    r = scene["particle_radius"]
    d = r * 2
    h = 4 * r
    ws = fluid["water_size"]
    particle_num = int(ws[0] / d * ws[1] / d * ws[2] / d)
    box = [scene["box_max"][k] - scene["box_min"][k] for k in range(3)]
    x_cnt = int(box[0] / d + 1)
    z_cnt = int(box[2] / d + 1)
    bottom = x_cnt * z_cnt
    one_round = x_cnt * z_cnt - (x_cnt - 2) * (z_cnt - 2)
    layer = int(math.ceil((box[1] - d) / d))
    boundary_num = layer * one_round + bottom * 2
    grid_num = tuple(int(math.ceil(box[k] / h)) + 1 for k in range(3))
    return particle_num, boundary_num, grid_num


def init_fluid_positions(config, particle_num, ids=None):
    """Fluid lattice of init_particle_pos (PS:142-151).  Returns (N, 3) float32.  `ids` (optional)
    restricts the result to a subset of the global lattice indices (multi-GPU slabs)."""
    scene, fluid = config["scene"], config["fluid"]
    r = scene["particle_radius"]
    d = r * 2
    ws, sp = fluid["water_size"], fluid["start_pos"]
    x_num_d = ws[0] / d
    z_num_d = ws[2] / d
    n = particle_num
    if n < (1 << 24):
        fi = (np.arange(n, dtype=np.int32) if ids is None else np.asarray(ids, dtype=np.int32)).astype(F32)
        x_num, z_num = F32(x_num_d), F32(z_num_d)
        xz_num = x_num * z_num                                     # PS:145: a product of two f32 kernel locals
        x = fi - x_num * np.floor(fi / x_num)                      # PS:147  i % x_num (float mod)
        t = np.floor(fi / x_num)
        z = t - z_num * np.floor(t / z_num)                        # PS:148
        y = (fi / xz_num).astype(np.int32).astype(F32)             # PS:149  int(i / xz_num)
    else:
        # the reference's f32 index arithmetic is inexact beyond 2^24: integer lattice
        i = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
        xi, zi = int(round(x_num_d)), int(round(z_num_d))
        x = (i % xi).astype(F32)
        z = ((i // xi) % zi).astype(F32)
        y = (i // (xi * zi)).astype(F32)
    rad = F32(r)
    pos = np.empty((x.shape[0], 3), dtype=F32)
    pos[:, 0] = (x * rad) * F32(2.0) + F32(sp[0])                  # PS:150
    pos[:, 1] = (y * rad) * F32(2.0) + F32(sp[1])
    pos[:, 2] = (z * rad) * F32(2.0) + F32(sp[2])
    return pos


def init_boundary_positions(config, boundary_num):
    """One-layer boundary shell of init_particle_pos (PS:155-195).  Returns (Nb, 3) float32."""
    scene = config["scene"]
    r = scene["particle_radius"]
    dd = r * 2
    box_x = scene["box_max"][0] - scene["box_min"][0]
    box_z = scene["box_max"][2] - scene["box_min"][2]
    # PS:155-158: `box` is a kernel local (f32), so these two counts are f32 arithmetic, unlike derive_sizes (host, fp64)
    x_cnt = int(F32(box_x) / F32(dd) + F32(1))
    z_cnt = int(F32(box_z) / F32(dd) + F32(1))
    xr, zr = x_cnt - 1, z_cnt - 1
    bottom = x_cnt * z_cnt
    one_round = x_cnt * z_cnt - (x_cnt - 2) * (z_cnt - 2)
    nb = boundary_num
    d = F32(dd)
    i = np.arange(nb, dtype=np.int64)
    x = np.zeros(nb, dtype=F32)
    y = np.zeros(nb, dtype=F32)
    z = np.zeros(nb, dtype=F32)

    m0 = i < bottom                                                # PS:164-168
    i0 = i[m0]
    x[m0] = (i0 % x_cnt).astype(F32) * d
    z[m0] = np.floor(i0.astype(F32) / F32(x_cnt)) * d

    m1 = (i >= bottom) & (i < nb - bottom)                         # PS:169-189
    idx = i[m1] - bottom
    layer = np.floor(idx.astype(F32) / F32(one_round)).astype(np.int64)
    y1 = d * (layer + 1).astype(F32)
    idx = idx - layer * one_round + 1
    x1 = np.zeros(idx.shape, dtype=F32)
    z1 = np.zeros(idx.shape, dtype=F32)
    a = idx <= xr
    x1[a] = (idx[a] % xr).astype(F32) * d
    b = (xr < idx) & (idx <= xr + zr)
    x1[b] = F32(xr) * d
    z1[b] = ((idx[b] - x_cnt) % zr).astype(F32) * d
    c = (xr + zr < idx) & (idx <= 2 * xr + zr)
    x1[c] = ((2 * xr + zr - idx[c]) % xr + 1).astype(F32) * d
    z1[c] = F32(zr) * d
    e = (2 * xr + zr < idx) & (idx <= 2 * (xr + zr))
    z1[e] = ((2 * (xr + zr) - idx[e]) % zr + 1).astype(F32) * d
    x[m1], y[m1], z[m1] = x1, y1, z1

    m2 = i >= nb - bottom                                          # PS:190-195
    i2 = i[m2] - (nb - bottom)
    x[m2] = (i2 % x_cnt).astype(F32) * d
    y[m2] = F32(scene["box_max"][1])
    z[m2] = (i2.astype(F32) / F32(x_cnt)).astype(np.int64).astype(F32) * d
    return np.stack([x, y, z], axis=1)


# ---- rigid body: mesh loading and voxelisation (PS:42-50; trimesh is not installed) ---------------

def load_mesh(path):
    """Minimal STL (binary/ascii) and OBJ reader -> (vertices (V,3) f64, faces (F,3) int)."""
    if not os.path.exists(path):
        # several shipped configs write ./obj/cube1.stl for obj/cube1.STL (SURVEY B-R5)
        d, f = os.path.split(path)
        for cand in os.listdir(d or "."):
            if cand.lower() == f.lower():
                path = os.path.join(d, cand)
                break
    ext = os.path.splitext(path)[1].lower()
    if ext == ".obj":
        vs, fs = [], []
        with open(path) as fh:
            for line in fh:
                p = line.split()
                if not p:
                    continue
                if p[0] == "v":
                    vs.append([float(p[1]), float(p[2]), float(p[3])])
                elif p[0] == "f":
                    ids = [int(t.split("/")[0]) - 1 for t in p[1:]]
                    for k in range(1, len(ids) - 1):
                        fs.append([ids[0], ids[k], ids[k + 1]])
        return np.asarray(vs, dtype=np.float64), np.asarray(fs, dtype=np.int64)
    with open(path, "rb") as fh:
        data = fh.read()
    ntri = struct.unpack_from("<I", data, 80)[0] if len(data) >= 84 else 0
    if len(data) == 84 + 50 * ntri:
        rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                            count=ntri, offset=84)
        tri = rec["v"].astype(np.float64).reshape(-1, 3)
    else:
        tri = np.asarray([[float(t) for t in ln.split()[1:4]] for ln in data.decode(errors="ignore").splitlines()
                          if ln.strip().startswith("vertex")], dtype=np.float64)
    verts, inv = np.unique(tri, axis=0, return_inverse=True)
    return verts, inv.reshape(-1, 3)


def voxelize(vertices, faces, pitch, fill=True):
    """Restatement of trimesh `mesh.voxelized(pitch)[.fill()].points` (subdivide method): subdivide
    every triangle until its edges are <= pitch/2, round vertices / pitch to integer voxel ids, fill
    interior holes, return the voxel centres id * pitch.  UNVERIFIED against trimesh (not installed)."""
    v = np.asarray(vertices, dtype=np.float64)
    tri = v[np.asarray(faces)]
    max_edge = pitch / 2.0
    pts = [v]
    for _ in range(32):
        e = np.stack([np.linalg.norm(tri[:, 0] - tri[:, 1], axis=1), np.linalg.norm(tri[:, 1] - tri[:, 2], axis=1),
                      np.linalg.norm(tri[:, 2] - tri[:, 0], axis=1)], axis=1)
        big = e.max(axis=1) > max_edge
        if not big.any():
            break
        t = tri[big]
        m01, m12, m20 = (t[:, 0] + t[:, 1]) / 2, (t[:, 1] + t[:, 2]) / 2, (t[:, 2] + t[:, 0]) / 2
        pts.append(np.concatenate([m01, m12, m20]))
        tri = np.concatenate([np.stack([t[:, 0], m01, m20], 1), np.stack([m01, t[:, 1], m12], 1),
                              np.stack([m20, m12, t[:, 2]], 1), np.stack([m01, m12, m20], 1)])
    allp = np.concatenate(pts)
    ids = np.unique(np.round(allp / pitch).astype(np.int64), axis=0)
    lo = ids.min(axis=0)
    dense = np.zeros(tuple(ids.max(axis=0) - lo + 1), dtype=bool)
    dense[tuple((ids - lo).T)] = True
    if fill:
        from scipy import ndimage
        dense = ndimage.binary_fill_holes(dense)
    out = np.argwhere(dense) + lo
    return (out * pitch).astype(np.float64)


def rigid_points_from_config(solid, base_dir="."):
    """Voxel points and mesh vertices of the 'solid' block (PS:42-57).  An explicit `points` entry
    (path to an .npy file or a nested list) overrides the voxeliser."""
    path = solid.get("mesh")
    if not os.path.isabs(path):
        path = os.path.join(base_dir, path)
    verts, faces = load_mesh(path)
    verts = verts * solid.get("scale", 1)                          # PS:43 apply_scale
    pts = solid.get("points")
    if pts is not None:
        pts = np.load(pts) if isinstance(pts, str) else np.asarray(pts)
    else:
        pts = voxelize(verts, faces, solid.get("voxel_radius") * 2, solid.get("fill", True))
    return pts.astype(F32), verts.astype(F32), faces


def rotation3d(ang_x, ang_y, ang_z):
    """ti.math.rotation3d as recalled in SURVEY App. A-11 (3x3 part), f32 arithmetic."""
    yaw, pitch, roll = F32(ang_z), F32(ang_x), F32(ang_y)
    ch, sh = np.cos(yaw, dtype=F32), np.sin(yaw, dtype=F32)
    cp, sp = np.cos(pitch, dtype=F32), np.sin(pitch, dtype=F32)
    cb, sb = np.cos(roll, dtype=F32), np.sin(roll, dtype=F32)
    return np.array([[ch * cb + sh * sp * sb, sb * cp, -sh * cb + ch * sp * sb],
                     [-ch * sb + sh * sp * cb, cb * cp, sb * sh + ch * sp * cb],
                     [sh * cp, -sp, ch * cp]], dtype=F32)
