"""The oracle (and, with -m gpu, the strict CUDA path) against the reference's OWN solver sources, executed.

tests/golden/refshim_<case>.npz were written by tests/golden/make_reference_shim_golden.py: the unmodified
ParticleSystem.py / solver_base.py / <name>_solver.py of the reference, imported from /root/reference and run under
tests/golden/ti_shim/taichi (a stand-in for the Taichi front end with Taichi's scalar rules; its docstring lists them).
Every array the reference's solver holds after each step() is compared BIT FOR BIT, together with the iteration counts and
residuals the reference prints.  What this pins: the transcription (statements, operand order, loop structure, constants,
quirks B-1 ... B-13).  What it cannot pin: Taichi's own code generation -- the stand-in follows SURVEY.md appendix A, the
same reading the oracle follows -- so oracle/sph_oracle.h keeps the words "parity unpinned" for the compiled reference."""
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
CASES = sorted(f[len("refshim_"):-len(".npz")] for f in os.listdir(GOLD) if f.startswith("refshim_") and f.endswith(".npz"))

sys.path.insert(0, GOLD)
import make_reference_shim_golden as gen  # noqa: E402  (case table and field lists; importing it runs nothing)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(bits(a), bits(b))


def describe(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return "shapes %s / %s" % (a.shape, b.shape)
    d = np.abs(a - b)
    k = int(np.argmax(d))
    return "max |diff| %.3e at flat index %d (%r vs %r), %d of %d differ" % (d.max(), k, a.reshape(-1)[k], b.reshape(-1)[k],
                                                                           int((d > 0).sum()), d.size)


ORACLE_NAME = {"constrain": "pbf_constrain", "delta_pos": "pbf_delta_pos", "constrain_derivative": "pbf_constrain_derivative"}
RIGID_FIELDS = {"pos": "rpos", "vel": "rvel", "acc": "racc", "force": "rforce", "omega": "romega", "alpha": "ralpha",
                "volume": "rvol", "mass": "rmass"}


def rigid_equal(o, d, tag, case):
    for ref_name, name in RIGID_FIELDS.items():
        ref = d["rigid_%s_%s" % (ref_name, tag)]
        assert same(o.field(name), ref), "%s, state %s, rigid %s: %s" % (case, tag, ref_name, describe(o.field(name), ref))
    for ref_name, name in (("centroid", "centroid"), ("inertia_inv", "inertia_inv"), ("vertices", "rverts")):
        ref = d["rigid_%s_%s" % (ref_name, tag)]
        got = o.field(name).reshape(ref.shape)
        assert same(got, ref), "%s, state %s, %s: %s" % (case, tag, ref_name, describe(got, ref))


def load(case):
    d = np.load(os.path.join(GOLD, "refshim_%s.npz" % case))
    cfg = json.loads(str(d["config_json"]))
    return d, cfg, str(d["solver"]), int(d["steps"])


def test_the_committed_cases_are_the_generators_cases():
    core = {"wcsph_block", "wcsph_clamp", "wcsph_tiny", "dfsph_block", "dfsph_lattice", "dfsph_clamp", "pcisph_block", "iisph_block"}
    assert core <= set(CASES) <= set(gen.CASES), "run tests/golden/make_reference_shim_golden.py (it needs /root/reference)"
    for case in CASES:
        d, cfg, solver, steps = load(case)
        want = gen.CASES[case][0]
        if isinstance(want, str):                       # one of the reference's own scene files
            if not os.path.isdir(REF):
                continue
            with open(os.path.join(REF, want[len("ref:"):])) as fh:
                want = json.load(fh)
        assert (cfg, steps) == (want, gen.CASES[case][1]), case
        # and the runs are not trivial: 150+ particles, the pressure loops iterate
        assert int(d["particle_num"]) >= 150 or case.endswith("_tiny"), case
    d = load("dfsph_block")[0]
    assert d["log_df_div_1"][0] >= 2 and d["log_df_den_1"][0] >= 2
    assert load("pcisph_block")[0]["log_pc_1"][0] >= 2 and load("iisph_block")[0]["log_ii_1"][0] >= 2


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_the_executed_reference_source(case):
    from oracle import oracle as O
    d, cfg, solver, steps = load(case)
    pts = verts = None
    if "solid" in cfg:        # the voxel points are this repository's restatement of trimesh (B-R1), the same on both sides
        from cfd_taichi_b200 import scene
        pts, verts, _ = scene.rigid_points_from_config(cfg["solid"], ROOT)
    o = O.Oracle(cfg, solver=solver, rigid_points=pts, rigid_vertices=verts, threads=1)
    if solver == "pbf":       # update_all_pos is ONE loop that moves particle i and then reads its neighbours: on one thread,
        o.set_scalar("pbf_update_mode", 1)    # neighbours j < i are already moved (the literal order; the CUDA path is Jacobi)
    # construction: ParticleSystem.__init__ (sizes, lattice, boundary shell, Akinci volumes), solver __init__
    assert int(o.scalar("particle_num")) == int(d["particle_num"])
    assert int(o.scalar("boundary_particles_num")) == int(d["boundary_particles_num"])
    assert [int(o.scalar("grid_" + a)) for a in "xyz"] == list(d["grid_num"])
    assert same(o.field("pos"), d["lattice_pos"]), describe(o.field("pos"), d["lattice_pos"])
    assert same(o.field("bpos"), d["boundary_pos"]), describe(o.field("bpos"), d["boundary_pos"])
    assert same(o.field("bvol"), d["boundary_volume"]), describe(o.field("bvol"), d["boundary_volume"])
    if solver == "pcisph":
        assert np.float32(o.scalar("pc_delta")) == d["pc_delta"], (o.scalar("pc_delta"), d["pc_delta"])
        assert o.scalar("pc_beta") == float(d["pc_beta"])
    if pts is not None:       # ParticleSystem.py:198-295: rotation, offset, Akinci volumes, mass, centroid, inertia and inverse
        assert same(o.field("inertia").reshape(9), d["rigid_inertia"]), describe(o.field("inertia").reshape(9), d["rigid_inertia"])
        o.field("rvel")[:] = d["rigid_vel_0"]
        rigid_equal(o, d, "0", case)
    o.field("pos")[:] = d["pos0"]
    o.field("vel")[:] = d["vel0"]
    coupled = 0.0
    for s in range(1, steps + 1):
        o.step(1, rigid=False)
        if pts is not None:   # main.py:166-171: the fluid step leaves the gathered fluid->rigid forces, then the rigid step
            ref = d["rigid_force_fluid_%d" % s]
            assert same(o.field("rforce"), ref), "%s, step %d, fluid->rigid force: %s" % (case, s, describe(o.field("rforce"), ref))
            coupled = max(coupled, float(np.abs(ref).max()))
            O.lib().orc_rigid_step(o._h)
            rigid_equal(o, d, str(s), case)
            for name, key in (("rs_omega", "rs_omega"), ("rs_attitude", "rs_attitude")):
                assert same(o.field(name).reshape(3), d["%s_%d" % (key, s)]), "%s, step %d, %s" % (case, s, name)
            assert np.float32(o.scalar("rs_dt")) == d["rs_dt_%d" % s]
        for name in ["pos", "vel", "cell3"] + gen.FIELDS[solver]:
            ref = d["%s_%d" % (name, s)]
            got = o.field(ORACLE_NAME.get(name, name))
            assert same(got, ref), "%s, step %d, %s: %s" % (case, s, name, describe(got, ref))
        assert np.float32(o.scalar("delta_time")) == d["delta_time_%d" % s]
        if solver == "dfsph":
            cnt, first, err = d["log_df_div_%d" % s]
            assert (int(o.scalar("df_div_iters")), np.float32(o.scalar("df_div_first_err")), np.float32(o.scalar("df_div_err"))) == (
                int(cnt), np.float32(first), np.float32(err))
            cnt, err = d["log_df_den_%d" % s]
            assert int(o.scalar("df_den_iters")) == int(cnt)
            assert np.float32(o.scalar("df_den_err")) == np.float32(err), (o.scalar("df_den_err"), err)
        if solver == "pcisph":
            cnt, err = d["log_pc_%d" % s]
            assert (int(o.scalar("pc_iters")), np.float32(o.scalar("pc_err"))) == (int(cnt), np.float32(err))
        if solver == "iisph":
            cnt, res = d["log_ii_%d" % s]
            assert (int(o.scalar("ii_iters")), np.float32(o.scalar("ii_residual"))) == (int(cnt), np.float32(res))
    if pts is not None:
        assert (coupled > 1.0) == bool(cfg["solver"].get("fs_couple", True)), "the case must couple (or, uncoupled, must not)"
    o.close()


# every case but the last one added (fs_couple: false), which no GPU run of this round has seen yet
GPU_CASES = [c for c in CASES if c != "wcsph_rigid_uncoupled"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GPU_CASES)
def test_cuda_strict_reproduces_the_executed_reference_source(built, case):
    """The same files against the product: strict kernels through the reference-named Python classes."""
    import torch
    from conftest import quiet_ps, quiet_solver
    d, cfg, solver, steps = load(case)
    if solver == "pbf":
        pytest.skip("the fixture holds the one-thread order of update_all_pos; the CUDA path's contract is the two-phase order "
                    "(oracle/sph_oracle_pbf.inc, tests/test_gpu_solvers.py)")
    ps = quiet_ps(cfg, strict=True, solver_name=solver)
    cls = getattr(importlib.import_module("cfd_taichi_b200.%s_solver" % solver), "%s_solver" % solver)
    sol = quiet_solver(cls, ps, cfg)
    n = ps.particle_num
    assert n == int(d["particle_num"]) and ps.boundary_particles_num == int(d["boundary_particles_num"])
    assert same(ps.fluid_particles.pos.to_numpy(), d["lattice_pos"])
    assert same(ps.boundary_particles.pos.to_numpy(), d["boundary_pos"])
    assert same(ps.boundary_particles.volume.to_numpy(), d["boundary_volume"])
    rs = None
    if "solid" in cfg:
        import ctypes
        from cfd_taichi_b200 import _lib
        from cfd_taichi_b200.rigid_solver import rigid_solver
        rs = rigid_solver(ps, cfg)
        nr = ps.rigid_particles_num
        assert same(ps.rigid_particles.pos.to_numpy(), d["rigid_pos_0"])
        assert same(ps.rigid_particles.volume.to_numpy(), d["rigid_volume_0"])
        info = ps.rigid_state()
        assert same(np.array(list(info.inertia_inv), dtype=np.float32), d["rigid_inertia_inv_0"])
        # the body's initial velocity: the reference holds it in rigid_particles.vel (rigid_solver.py:43 reads element 0),
        # the library in its device-side body state as well
        v0 = d["rigid_vel_0"][0]
        ps._rvel4[:nr, :3] = torch.from_numpy(v0).to(ps._device)
        for k in range(3):
            info.vel[k] = float(v0[k])
        # ... together with the largest surface speed of the body, which the library caches at the end of every rigid step
        # for the next step's adaptive time step (dfsph_solver.py:104-111 recomputes it from rigid_particles every step)
        info.max_surface_vel = float(np.sqrt((v0[0] * v0[0] + v0[1] * v0[1]) + v0[2] * v0[2]))
        _lib.check(ps._lib.sph_rigid_set_state(ps._h, ctypes.byref(info)), ps._h)
    ps._pos4[:n, :3] = torch.from_numpy(d["pos0"]).to(ps._device)
    ps._vel4[:n, :3] = torch.from_numpy(d["vel0"]).to(ps._device)
    for s in range(1, steps + 1):
        sol.step()
        if rs is not None:
            assert same(ps.rigid_particles.force.to_numpy(), d["rigid_force_fluid_%d" % s]), "%s, step %d, fluid->rigid force" % (case, s)
            rs.step()
            for name in ("pos", "vel"):
                ref = d["rigid_%s_%d" % (name, s)]
                got = getattr(ps.rigid_particles, name).to_numpy()
                assert same(got, ref), "%s, step %d, rigid %s: %s" % (case, s, name, describe(got, ref))
            info = ps.rigid_state()
            assert same(np.array(list(info.omega), dtype=np.float32), d["rs_omega_%d" % s])
            assert same(np.array(list(info.centroid), dtype=np.float32), d["rigid_centroid_%d" % s])
        got = {"pos": ps.fluid_particles.pos.to_numpy(), "vel": ps.fluid_particles.vel.to_numpy(), "rho": sol.rho.to_numpy()}
        for name, a in got.items():
            ref = d["%s_%d" % (name, s)]
            assert same(a, ref), "%s, step %d, %s: %s" % (case, s, name, describe(a, ref))
        st = sol.stats()
        assert np.float32(st.delta_time) == d["delta_time_%d" % s]
        if solver == "dfsph":
            assert (st.div_iters, st.den_iters) == (int(d["log_df_div_%d" % s][0]), int(d["log_df_den_%d" % s][0]))
        if solver == "pcisph":
            assert st.pc_iters == int(d["log_pc_%d" % s][0])
        if solver == "iisph":
            assert st.ii_iters == int(d["log_ii_%d" % s][0])
        assert st.error_flags == 0
    ps.close()


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference sources are not on this machine")
def test_fixture_regenerates_from_the_reference_source(tmp_path):
    """Re-run the smallest case from /root/reference now (a subprocess with a clean module table) and require the committed
    file byte-equal array by array: the fixtures are what the generator produces from the reference as it lies there."""
    out = tmp_path / "again.npz"
    r = subprocess.run([sys.executable, os.path.join(GOLD, "make_reference_shim_golden.py"), "wcsph_tiny", "--out", str(out)],
                       capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-3000:]
    a, b = np.load(out), np.load(os.path.join(GOLD, "refshim_wcsph_tiny.npz"))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert same(a[k], b[k]) or (a[k].dtype.kind == "U" and str(a[k]) == str(b[k])), k


def test_colour_maps_of_the_executed_reference():
    """solver_base.visualize_rho / visualize_neighbour (SB:219-245) as the reference computes them, against the statement of
    them that tests/test_gpu_dfsph.py holds the CUDA colour maps to: b = (x - min) / (max - min) in f32, rgb = (0, 0.28, b);
    the neighbour count is get_neighbour_count's (quirk B-7), i.e. the oracle's."""
    from oracle import oracle as O
    d, cfg, solver, steps = load("wcsph_visualize")
    o = O.Oracle(cfg, solver=solver, threads=1)
    o.field("pos")[:] = d["pos0"]
    o.field("vel")[:] = d["vel0"]
    o.step(steps, rigid=False)
    assert same(o.field("pos"), d["pos_%d" % steps])
    o.phase("reset_grid_update_grid")
    o.phase("neighbour_counts")
    for key, x in (("rgb_rho", o.field("rho").copy()), ("rgb_neighbour", o.field("nbr_count").astype(np.float32))):
        rgb = d[key]
        b = (x - x.min()) / (x.max() - x.min())
        assert b.dtype == np.float32 and 0.0 == b.min() and b.max() == 1.0
        assert same(rgb[:, 2], b), "%s: %s" % (key, describe(rgb[:, 2], b))
        assert same(rgb[:, 0], np.zeros_like(b)) and same(rgb[:, 1], np.full_like(b, np.float32(0.28)))
    o.close()


@pytest.mark.parametrize("case", CASES)
def test_product_host_formulas_against_the_executed_reference(case):
    """The product's own host statement of ParticleSystem.__init__ / init_particle_pos (cfd_taichi_b200/scene.py: the sizes
    every device array is allocated with, the lattice and the boundary shell the CUDA initialisers are tested against) on the
    executed reference's output -- no oracle involved."""
    from cfd_taichi_b200 import scene
    d, cfg, solver, steps = load(case)
    pn, bn, grid = scene.derive_sizes(cfg)
    assert (pn, bn, list(grid)) == (int(d["particle_num"]), int(d["boundary_particles_num"]), list(d["grid_num"]))
    assert same(scene.init_fluid_positions(cfg, pn), d["lattice_pos"])
    got = scene.init_boundary_positions(cfg, bn)
    assert same(got, d["boundary_pos"]), describe(got, d["boundary_pos"])
