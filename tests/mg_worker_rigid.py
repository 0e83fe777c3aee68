"""torchrun worker of tests/test_gpu_multigpu.py: DFSPH with a rigid box over x-slabs.  The body is replicated
on every rank, the fluid->rigid forces are partial sums over the ranks' owned fluid particles (one all-reduce
per rigid step), so the comparison with the single-domain run is by tolerance, not bit for bit."""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cfd_taichi_b200 import scene, scenes  # noqa: E402
from cfd_taichi_b200.ParticleSystem import ParticleSystem  # noqa: E402
from cfd_taichi_b200.dfsph_solver import dfsph_solver  # noqa: E402
from cfd_taichi_b200.rigid_solver import rigid_solver  # noqa: E402


def box_points(lo, hi, pitch=0.05):
    ax = [np.arange(int(round(lo[k] / pitch)), int(round(hi[k] / pitch)) + 1) * pitch for k in range(3)]
    return np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # the fluid block touches the body from the first step on, and the cut between the slabs runs through both
    cfg = scenes.make_scene([2.0, 2.0, 1.0], [0.1, 0.1, 0.1], [1.0, 0.8, 0.8], "dfsph", 1e-4,
                            solid={"mesh": "unused", "voxel_radius": 0.025, "rho_0": 2000, "scale": 1,
                                   "pos_offset": [1.12, 0.0, 0.2], "attitude_offset": [0.0, 0.0, 0.0],
                                   "fill": True, "active": True})
    pts = box_points([0, 0, 0], [0.3, 0.4, 0.5])
    verts = np.array([[x, y, z] for x in (0.0, 0.3) for y in (0.0, 0.4) for z in (0.0, 0.5)], dtype=np.float32)
    scene.rigid_points_from_config = lambda solid, base_dir=".": (pts, verts, None)

    def run(slab):
        with contextlib.redirect_stdout(io.StringIO()):
            ps = ParticleSystem(cfg, strict=True, solver_name="dfsph", slab=slab)
            sol = dfsph_solver(ps, cfg)
            rs = rigid_solver(ps, cfg)
        for _ in range(steps):
            sol.step()
            rs.step()
        return ps, sol, rs

    ps, sol, rs = run((rank, world))
    gid, pos, v4 = ps.owned_state()
    info = ps.rigid_state()
    st = sol.stats()
    out = dict(gid=gid, pos=pos, vel=v4, cen=list(info.centroid), rvel=list(info.vel), omega=list(info.omega),
               div=st.div_iters, den=st.den_iters, flags=st.error_flags, rank_cut=ps._slab["cuts"])
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(out, gathered, dst=0)
    ok = True
    if rank == 0:
        ps1, sol1, rs1 = run(None)
        n = ps1.particle_num
        ref_pos = ps1._pos4[:n, :3].cpu().numpy()
        info1 = ps1.rigid_state()
        gids = np.concatenate([g["gid"] for g in gathered])
        pos = np.concatenate([g["pos"] for g in gathered])[np.argsort(gids)]
        perm_ok = np.array_equal(np.sort(gids), np.arange(n))
        dpos = float(np.abs(pos - ref_pos).max())
        dcen = float(np.abs(np.array(gathered[0]["cen"]) - np.array(list(info1.centroid))).max())
        drv = float(np.abs(np.array(gathered[0]["rvel"]) - np.array(list(info1.vel))).max())
        same = all(g["cen"] == gathered[0]["cen"] and g["omega"] == gathered[0]["omega"] for g in gathered)
        moved = float(np.abs(np.array(list(info1.vel))).max())
        print("MGRIGID perm_ok=%s replicas_identical=%s dpos=%.3e dcen=%.3e drvel=%.3e rigid_speed=%.3e cuts=%s flags=%s" % (
            perm_ok, same, dpos, dcen, drv, moved, gathered[0]["rank_cut"], [g["flags"] for g in gathered]), flush=True)
        ok = perm_ok and same and dpos <= 1e-4 and dcen <= 1e-5 and moved > 0 and all(g["flags"] == 0 for g in gathered)
        ps1.close()
    ps.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
