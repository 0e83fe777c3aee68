"""GPU parity: rigid-fluid coupling (BASELINE.json configs[3]: the dam_flush_cube scene with DFSPH)
against the CPU oracle: Akinci rigid volumes, mass properties, coupled DFSPH sweeps, gathered
fluid->rigid forces, rigid-body step (force / torque reductions, rotation, wall contact)."""
import numpy as np
import pytest

from cfd_taichi_b200 import scene, scenes
from cfd_taichi_b200.dfsph_solver import dfsph_solver
from cfd_taichi_b200.rigid_solver import rigid_solver
from conftest import quiet_ps, quiet_solver
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def box_points(lo, hi, pitch=0.05):
    """Voxel centres of an axis-aligned box (what the voxeliser returns for cube1.STL, SURVEY B-R1)."""
    ax = [np.arange(int(round(lo[k] / pitch)), int(round(hi[k] / pitch)) + 1) * pitch for k in range(3)]
    g = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3)
    return g.astype(np.float32)


def rigid_scene(small=True):
    if small:
        cfg = scenes.make_scene([2.0, 2.0, 1.0], [0.1, 0.1, 0.1], [0.6, 0.8, 0.8], "dfsph", 1e-4,
                                solid={"mesh": "unused", "voxel_radius": 0.025, "rho_0": 2000, "scale": 1,
                                       "pos_offset": [0.85, 0.0, 0.2], "attitude_offset": [0.0, 0.0, 0.0],
                                       "fill": True, "active": True})
        pts = box_points([0, 0, 0], [0.3, 0.4, 0.5])
    else:
        cfg = scenes.shipped("dam_flush_cube", "dfsph")
        pts = box_points([0, 0, 0], [0.8, 0.5, 1.0])
    verts = np.array([[x, y, z] for x in (pts[:, 0].min(), pts[:, 0].max()) for y in (pts[:, 1].min(), pts[:, 1].max())
                      for z in (pts[:, 2].min(), pts[:, 2].max())], dtype=np.float32)
    return cfg, pts, verts


def make(cfg, pts, verts, strict, monkeypatch):
    monkeypatch.setattr(scene, "rigid_points_from_config", lambda solid, base_dir=".": (pts, verts, None))
    ps = quiet_ps(cfg, strict=strict, solver_name="dfsph")
    sol = quiet_solver(dfsph_solver, ps, cfg)
    rs = rigid_solver(ps, cfg)
    o = O.Oracle(cfg, solver="dfsph", rigid_points=pts, rigid_vertices=verts, threads=1)
    return ps, sol, rs, o


def test_rigid_init_bit_exact(built, monkeypatch):
    cfg, pts, verts = rigid_scene()
    ps, sol, rs, o = make(cfg, pts, verts, True, monkeypatch)
    assert ps.rigid_particles_num == len(pts)
    assert np.array_equal(ps.rigid_particles.pos.to_numpy(), o.field("rpos"))
    assert np.array_equal(ps.rigid_particles.volume.to_numpy(), o.field("rvol"))
    assert np.array_equal(ps.rigid_particles.mass.to_numpy(), o.field("rmass"))
    info = ps.rigid_state()
    assert np.array_equal(np.array(list(info.centroid), dtype=np.float32), o.field("centroid").reshape(-1))
    assert np.array_equal(np.array(list(info.inertia), dtype=np.float32), o.field("inertia").reshape(-1))
    assert np.array_equal(np.array(list(info.inertia_inv), dtype=np.float32), o.field("inertia_inv").reshape(-1))
    ps.close(); o.close()


def test_coupled_steps_strict_bit_exact(built, monkeypatch):
    cfg, pts, verts = rigid_scene()
    ps, sol, rs, o = make(cfg, pts, verts, True, monkeypatch)
    for step in range(6):
        sol.step()
        # fluid->rigid forces gathered during the density solve, before the rigid step consumes them
        assert np.array_equal(ps.rigid_particles.force.to_numpy(), o_force_after_fluid(o)), "step %d forces" % step
        rs.step()
        st = sol.stats()
        assert st.error_flags == 0
        assert (st.div_iters, st.den_iters) == (int(o.scalar("df_div_iters")), int(o.scalar("df_den_iters")))
        assert np.array_equal(ps.fluid_particles.pos.to_numpy(), o.field("pos")), "step %d fluid pos" % step
        assert np.array_equal(ps.fluid_particles.vel.to_numpy(), o.field("vel")), "step %d fluid vel" % step
        assert np.array_equal(ps.neighbour_counts().cpu().numpy(), o.field("nbr_count")), "step %d counts" % step
        info = ps.rigid_state()
        assert np.allclose(np.array(list(info.centroid)), o.field("centroid").reshape(-1), rtol=1e-6, atol=1e-7)
        assert np.allclose(ps.rigid_particles.pos.to_numpy(), o.field("rpos"), rtol=1e-6, atol=1e-7)
        assert np.array_equal(ps.rigid_particles.pos.to_numpy(), o.field("rpos")), "step %d rigid pos" % step
        assert np.array_equal(np.array(list(info.omega), dtype=np.float32), o.field("rs_omega").reshape(-1))
    ps.close(); o.close()


def o_force_after_fluid(o):
    """Advance the oracle by the fluid step only and return the accumulated rigid forces, then run its
    rigid step (main.py:166-171 ordering)."""
    import ctypes
    from oracle import oracle as orc
    orc.lib().orc_step(o._h)
    f = o.field("rforce").copy()
    orc.lib().orc_rigid_step(o._h)
    return f


def test_coupled_steps_fast_within_tolerance(built, monkeypatch):
    """Fast kernels on the coupled scene, every sweep in isolation (rigid forces included) at 1e-5 and identical loop
    decisions; the strict walk it is compared with must end bit-exact on the oracle.  The shipped dam_flush_cube
    scene gets the same treatment in tests/test_gpu_fast_parity.py."""
    from cfd_taichi_b200 import selfcheck
    cfg, pts, verts = rigid_scene()
    ps_s, sol_s, rs_s, o = make(cfg, pts, verts, True, monkeypatch)
    ps_f, sol_f, rs_f, o2 = make(cfg, pts, verts, False, monkeypatch)
    o2.close()
    for step in range(3):
        selfcheck.copy_caller_state(ps_f, sol_f, ps_s, sol_s)
        err, info = selfcheck.sweeps("dfsph", ps_s, sol_s, ps_f, sol_f, rigid=True)
        f_ref = o_force_after_fluid(o)
        assert np.array_equal(ps_s.rigid_particles.force.to_numpy(), f_ref)
        assert selfcheck.relinf(ps_f.rigid_particles.force.to_numpy(), f_ref) <= 1e-5
        rs_s.step(); rs_f.step()
        assert np.array_equal(ps_s.fluid_particles.pos.to_numpy(), o.field("pos")), "step %d" % step
        assert info["loop_flags_equal"] and info["neighbour_counts_equal"] and info["error_flags"] == (0, 0)
        assert info["iters"]["fast"] == info["iters"]["strict"] == (int(o.scalar("df_div_iters")), int(o.scalar("df_den_iters")))
        w, where = selfcheck.worst(err)
        assert w <= 1e-5, "step %d: %s off by %.3e" % (step, where, w)
        cf, cs = np.array(list(ps_f.rigid_state().centroid)), o.field("centroid").reshape(-1)
        assert selfcheck.relinf(cf, cs) <= 1e-5
    ps_s.close(); ps_f.close(); o.close()


def test_dam_flush_cube_scene(built, monkeypatch):
    """BASELINE.json configs[3] at full size: 56 447 fluid + 21 602 boundary + 3 927 rigid particles."""
    cfg, pts, verts = rigid_scene(small=False)
    ps, sol, rs, o = make(cfg, pts, verts, True, monkeypatch)
    assert (ps.particle_num, ps.boundary_particles_num, ps.rigid_particles_num) == (56447, 21602, 3927)
    for step in range(2):
        sol.step()
        f_ref = o_force_after_fluid(o)
        assert np.array_equal(ps.rigid_particles.force.to_numpy(), f_ref)
        rs.step()
        assert np.array_equal(ps.fluid_particles.pos.to_numpy(), o.field("pos"))
        assert np.array_equal(ps.rigid_particles.pos.to_numpy(), o.field("rpos"))
    ps.close(); o.close()


@pytest.mark.parametrize("solver", ["wcsph", "pcisph", "iisph"])
def test_other_solvers_coupled_strict_bit_exact(built, monkeypatch, solver):
    """The other three scatter sites (WC:126, PC:186, II:159) in gather form, plus the rigid terms of every
    sweep of those solvers, against the oracle."""
    from cfd_taichi_b200.iisph_solver import iisph_solver
    from cfd_taichi_b200.pcisph_solver import pcisph_solver
    from cfd_taichi_b200.wcsph_solver import wcsph_solver
    cls = {"wcsph": wcsph_solver, "pcisph": pcisph_solver, "iisph": iisph_solver}[solver]
    cfg, pts, verts = rigid_scene()
    cfg["solver"]["name"] = solver
    cfg["solver"]["delta_time"] = {"wcsph": 2.5e-4, "pcisph": 1e-4, "iisph": 2.5e-4}[solver]
    monkeypatch.setattr(scene, "rigid_points_from_config", lambda solid, base_dir=".": (pts, verts, None))
    ps = quiet_ps(cfg, strict=True, solver_name=solver)
    sol = quiet_solver(cls, ps, cfg)
    rs = rigid_solver(ps, cfg)
    o = O.Oracle(cfg, solver=solver, rigid_points=pts, rigid_vertices=verts, threads=1)
    if solver == "pcisph":
        assert sol.stats().pc_max_index == int(o.scalar("pc_max_index"))
        assert np.float32(sol.delta[None]) == np.float32(o.scalar("pc_delta"))
    for step in range(4):
        sol.step()
        f_ref = o_force_after_fluid(o)
        assert np.array_equal(ps.rigid_particles.force.to_numpy(), f_ref), "step %d rigid forces" % step
        rs.step()
        assert sol.stats().error_flags == 0
        assert np.array_equal(ps.fluid_particles.pos.to_numpy(), o.field("pos")), "step %d fluid pos" % step
        assert np.array_equal(ps.fluid_particles.vel.to_numpy(), o.field("vel")), "step %d fluid vel" % step
        assert np.array_equal(ps.rigid_particles.pos.to_numpy(), o.field("rpos")), "step %d rigid pos" % step
    ps.close(); o.close()


def test_rigid_scene_restart_is_bit_exact(built, monkeypatch, tmp_path):
    # a dam hitting the box, dumped in the middle of the contact and resumed in a fresh ParticleSystem
    from cfd_taichi_b200 import main as app
    cfg, pts, verts = rigid_scene()
    cfg["solid"]["pos_offset"] = [0.72, 0.0, 0.2]      # touching the fluid block from the first step on
    monkeypatch.setattr(scene, "rigid_points_from_config", lambda solid, base_dir=".": (pts, verts, None))

    def fresh():
        ps = quiet_ps(cfg, strict=True, solver_name="dfsph")
        return ps, quiet_solver(dfsph_solver, ps, cfg), rigid_solver(ps, cfg)

    ps, sol, rs = fresh()
    for _ in range(6):
        sol.step(); rs.step()
    app.save_state(str(tmp_path / "dump"), ps, sol, rs)
    for _ in range(6):
        sol.step(); rs.step()
    want = (ps.fluid_particles.pos.to_numpy(), ps._vel4[:ps.particle_num].cpu().numpy(), ps.rigid_particles.pos.to_numpy(),
            list(ps.rigid_state().centroid), list(ps.rigid_state().omega))
    assert any(abs(w) > 0 for w in want[4]) or np.abs(ps.rigid_particles.vel.to_numpy()).max() > 0
    ps.close()
    ps2, sol2, rs2 = fresh()
    app.load_state(str(tmp_path / "dump"), ps2, sol2)
    for _ in range(6):
        sol2.step(); rs2.step()
    got = (ps2.fluid_particles.pos.to_numpy(), ps2._vel4[:ps2.particle_num].cpu().numpy(), ps2.rigid_particles.pos.to_numpy(),
           list(ps2.rigid_state().centroid), list(ps2.rigid_state().omega))
    for a, b in zip(got[:3], want[:3]):
        assert np.array_equal(a, b)
    assert got[3] == want[3] and got[4] == want[4]
    ps2.close()
