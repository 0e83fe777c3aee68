"""Golden vectors from the reference's OWN solver sources, executed here.

Taichi cannot be installed in this container (no wheel, no network), so the reference cannot run as it is.  Its
solver sources, however, are plain Python once `import taichi` resolves: tests/golden/ti_shim/taichi is a stand-in that
rewrites every `@ti.kernel` / `@ti.func` the way Taichi's front end does and runs it with Taichi's scalar rules
(binary32 / int32 runtime values, compile-time Python constants, declared local types, by-reference template arguments,
array-of-structures fields, dynamic cell lists; see its docstring).  This script imports the UNMODIFIED files from
/root/reference (ParticleSystem.py, solver_base.py, <name>_solver.py), builds small scenes through the reference's own
constructors, sets a seeded perturbed state through the reference's own `from_numpy`, calls the reference's own `step()`
and writes the state after every step to tests/golden/refshim_<case>.npz.

    python tests/golden/make_reference_shim_golden.py            # every case (minutes: pure Python)
    python tests/golden/make_reference_shim_golden.py wcsph_block --out /tmp/x.npz

tests/test_reference_shim.py holds the oracle (and, with -m gpu, the strict CUDA path) to these files bit for bit.  What the files
pin: the oracle's transcription of the reference's statements, operand order, loop structure, constants and quirks.
What they cannot pin: Taichi's own code generation (SURVEY.md appendix A is still this repository's reading of it), which
is why the oracle's header keeps the words "parity unpinned" for the compiled reference.
"""
import contextlib
import io
import json
import os
import re
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SPH_REFERENCE_DIR", "/root/reference")


def block_scene(solver, dt, water=(0.35, 0.4, 0.35), box=(0.7, 0.8, 0.7), boundary_handle=True):
    return {"scene": {"box_min": [0.0, 0.0, 0.0], "box_max": list(box), "particle_radius": 0.025, "gravity": 9.8},
            "fluid": {"start_pos": [0.15, 0.1, 0.15], "water_size": list(water)},
            "solver": {"name": solver, "delta_time": dt, "iter_cnt": 1, "boundary_handle": boundary_handle}}


def other_radius_scene():
    cfg = block_scene("dfsph", 5e-4, water=(0.24, 0.24, 0.2), box=(0.56, 0.6, 0.5))
    cfg["scene"]["particle_radius"] = 0.02
    cfg["scene"]["gravity"] = 3.7
    cfg["fluid"]["start_pos"] = [0.12, 0.08, 0.12]
    return cfg


def rigid_scene(solver, dt, pos_offset, attitude_deg, scale=0.4, fs_couple=True):
    cfg = block_scene(solver, dt, water=(0.3, 0.3, 0.3), box=(0.8, 0.8, 0.8))
    cfg["fluid"]["start_pos"] = [0.1, 0.1, 0.1]
    cfg["solid"] = {"active": True, "attitude_offset": list(attitude_deg), "fill": True, "mesh": "./obj/cube1.STL",
                    "pos_offset": list(pos_offset), "rho_0": 2000, "scale": scale, "voxel_radius": 0.025}
    if not fs_couple:
        cfg["solver"]["fs_couple"] = False
    return cfg


# name -> (config, steps, perturbation or None[, initial velocity of the rigid body]); perturbation = (seed, position jitter in particle diameters, velocity scale,
# compression of the block about its centre: the rest lattice has density 682 < rho_0 (quirk B-3), so without compression no
# pressure solver would iterate)
CASES = {
    "wcsph_block": (block_scene("wcsph", 2.5e-4), 3, (11, 0.18, 1.5, 0.85)),
    "dfsph_block": (block_scene("dfsph", 1e-3), 2, (12, 0.18, 0.3, 0.85)),
    "dfsph_lattice": (block_scene("dfsph", 1e-3, water=(0.3, 0.3, 0.3)), 2, None),
    "dfsph_clamp": (block_scene("dfsph", 1e-3, water=(0.3, 0.3, 0.3), boundary_handle=False), 2, (13, 0.18, 2.5, 1.0)),
    "pcisph_block": (block_scene("pcisph", 1.5e-4, water=(0.3, 0.3, 0.3)), 2, (14, 0.1, 0.5, 0.86)),
    "pcisph_converging": (block_scene("pcisph", 1.5e-4, water=(0.3, 0.3, 0.3)), 2, (21, 0.1, 0.5, 0.93)),   # 25 iterations, then below 0.1 %
    "iisph_block": (block_scene("iisph", 2.5e-4, water=(0.3, 0.3, 0.3)), 2, (15, 0.1, 0.5, 0.86)),
    "wcsph_clamp": (block_scene("wcsph", 2.5e-4, water=(0.3, 0.3, 0.3), boundary_handle=False), 2, (16, 0.18, 2.5, 1.0)),
    # fluid-rigid coupling (ParticleSystem.py:198-307, rigid_solver.py, the material_solid branches of every sweep): a
    # 0.32 x 0.2 x 0.4 box of 315 rigid particles 0.07 above the block, pushed sideways and down so that it meets the fluid
    # and, in the second case, the floor (rigid_solver.kinematic's contact impulse)
    # (140 rigid particles here: get_neighbour_count reads fluid_particles.pos[rigid-local index], quirk B-7, which is
    # only defined while there are fewer rigid than fluid particles)
    "dfsph_rigid": (rigid_scene("dfsph", 1e-3, [0.12, 0.36, 0.1], [0.0, 0.0, 0.0], scale=0.3), 3, (18, 0.1, 0.5, 0.82), [0.3, -2.0, 0.1]),
    # (PCISPH adds the pressure force to the body once per pressure ITERATION, PC:186 inside the loop, and kinematic() clears
    # it once per step -- quirk B-R3 -- so 45 iterations leave 45 times the force: the body leaves at 80 m/s.  That is the
    # reference's behaviour, and it is what the fixture holds.)
    "pcisph_rigid": (rigid_scene("pcisph", 1.5e-4, [0.12, 0.385, 0.1], [0.0, 0.0, 0.0], scale=0.3), 2, (25, 0.1, 0.5, 0.97), [0.2, -1.0, 0.1]),
    "iisph_rigid": (rigid_scene("iisph", 2.5e-4, [0.12, 0.36, 0.1], [5.0, 0.0, 10.0], scale=0.3), 3, (26, 0.1, 0.5, 0.84), [0.2, -1.0, 0.1]),
    "wcsph_rigid_uncoupled": (rigid_scene("wcsph", 2.5e-4, [0.12, 0.36, 0.1], [0.0, 0.0, 0.0], scale=0.3, fs_couple=False), 2,
                              (27, 0.1, 0.5, 0.84), [0.2, -1.0, 0.1]),           # `fs_couple: false`: the body is in the grid, the sweeps skip it
    "wcsph_rigid_floor": (rigid_scene("wcsph", 2.5e-4, [0.42, 0.0512, 0.2], [0.0, 15.0, 0.0]), 4, (19, 0.1, 0.5, 0.9), [-0.5, -3.0, 0.2]),
    # PBF: pbf_solver.py cannot compile at the reference's HEAD (its tasks take integer (i, j), for_all_neighbor passes
    # structs: quirk B-14).  The ONE line that makes it run is the reference's own commented-out alternative at
    # ParticleSystem.py:468, `ret += task(i, neighbor_index)`, substituted in the function's source at import time
    # (index_based_for_all_neighbor below); everything else is executed as it is written
    "pbf_block": (block_scene("pbf", 2.5e-4, water=(0.3, 0.3, 0.3)), 2, (23, 0.12, 0.8, 0.86)),
    # the colour maps of solver_base.visualize_rho / visualize_neighbour (SB:219-245) after one step and a grid rebuild
    "wcsph_visualize": (block_scene("wcsph", 2.5e-4, water=(0.3, 0.3, 0.3)), 1, (24, 0.15, 1.0, 0.9)),
    # another particle radius and gravity than every shipped scene has (h = 0.08: the cull threshold, the kernel constants
    # and the boundary spacing all move)
    "dfsph_radius_002": (other_radius_scene(), 2, (22, 0.15, 0.8, 0.86)),
    # seconds, not minutes: the case tests/test_reference_shim.py re-runs from /root/reference on every CPU test run
    # the reference's own shipped scene files, as they are, from the lattice start (5 879 fluid + 9 002 boundary particles:
    # tens of minutes of pure Python per DFSPH step)
    "shipped_wcsph": ("ref:config/wcsph_config_backup.json", 2, None),
    "shipped_dfsph": ("ref:config/dfsph_config_backup.json", 3, None),
    "shipped_pcisph": ("ref:config/pcisph_config_backup.json", 1, None),       # pre_compute's delta on the reference's own scene
    "shipped_iisph": ("ref:config/iisph_config_backup.json", 1, None),
    "wcsph_tiny": (block_scene("wcsph", 2.5e-4, water=(0.2, 0.25, 0.2), box=(0.5, 0.5, 0.5)), 2, (17, 0.18, 1.5, 0.85)),
}

def index_based_for_all_neighbor(ps_mod, ti):
    """SURVEY.md 8(f) rank 4, "PBF with index-based semantics": activate the reference's own commented-out line PS:468."""
    import inspect
    import linecache
    import textwrap
    orig = ps_mod.ParticleSystem.for_all_neighbor.__ti_original__
    src = textwrap.dedent(inspect.getsource(orig))
    struct_call, index_call = "ret += task(particle, particle_j)", "ret += task(i, neighbor_index)"
    assert src.count(struct_call) == 1 and src.count("# " + index_call) == 1
    src = src.replace("# " + index_call, "#").replace(struct_call, index_call)
    src = "\n".join(l for l in src.splitlines() if not l.startswith("@")) + "\n"
    fname = "<%s: for_all_neighbor with line 468 active>" % os.path.join(REF, "ParticleSystem.py")
    linecache.cache[fname] = (len(src), None, src.splitlines(True), fname)
    ns = {}
    exec(compile(src, fname, "exec"), ps_mod.__dict__, ns)
    ps_mod.ParticleSystem.for_all_neighbor = ti.func(ns["for_all_neighbor"])


FIELDS = {
    "wcsph": ["rho", "pressure", "pressure_gradient", "boundary_acc", "viscosity", "tension"],
    "dfsph": ["rho", "alpha", "rho_adv", "rho_derivative", "vel_adv", "vel_adv_delta", "force_ext", "warm_start_k", "viscosity",
              "tension"],
    "pcisph": ["rho", "pos_predict", "vel_predict", "ext_force", "press_force", "rho_predict", "rho_err", "press_iter",
               "viscosity", "tension"],
    "pbf": ["rho", "constrain", "pos_predict", "delta_pos", "constrain_derivative", "pbf_lambda"],
    "iisph": ["rho", "v_adv", "f_adv", "d_ii", "a_ii", "d_ij", "rho_adv", "p_iter", "p_past", "r_sum", "f_press", "viscosity",
              "tension"],
}

LOG_PATTERNS = {
    "df_div": re.compile(r"\[divergence iteration\] count: (\S+), first error (\S+), error (\S+)"),
    "df_den": re.compile(r"\[density iteration\] count: (\S+), error (\S+)"),
    "pc": re.compile(r"\t\tIter cnt: (\S+), error: (\S+)"),
    "ii": re.compile(r"Iter cnt:  (\S+) (\S+)"),
}


def perturbed_state(lattice, seed, jitter, vscale, compress, diameter=0.05):
    rng = np.random.default_rng(seed)
    n = lattice.shape[0]
    centre = lattice.astype(np.float64).mean(axis=0)
    pos = centre + (lattice.astype(np.float64) - centre) * compress + rng.uniform(-jitter, jitter, size=(n, 3)) * diameter
    vel = rng.standard_normal(size=(n, 3)) * vscale
    return pos.astype(np.float32), vel.astype(np.float32)


def run_case(name):
    cfg, steps, pert = CASES[name][:3]
    rigid_vel = CASES[name][3] if len(CASES[name]) > 3 else None
    if isinstance(cfg, str):
        with open(os.path.join(REF, cfg[len("ref:"):])) as fh:
            cfg = json.load(fh)
    solver = cfg["solver"]["name"]
    for m in ("taichi", "trimesh", "ParticleSystem", "solver_base", "rigid_solver", solver + "_solver"):
        sys.modules.pop(m, None)
    sys.path[:0] = [os.path.join(HERE, "ti_shim"), REF]
    try:
        import importlib
        ps_mod = importlib.import_module("ParticleSystem")
        assert os.path.dirname(os.path.abspath(ps_mod.__file__)) == os.path.abspath(REF), ps_mod.__file__
        sol_mod = importlib.import_module(solver + "_solver")
        assert os.path.dirname(os.path.abspath(sol_mod.__file__)) == os.path.abspath(REF), sol_mod.__file__
        rig_mod = importlib.import_module("rigid_solver") if "solid" in cfg else None
        ti = importlib.import_module("taichi")
        if solver == "pbf":
            index_based_for_all_neighbor(ps_mod, ti)
    finally:
        del sys.path[:2]
    out = {"config_json": np.array(json.dumps(cfg)), "solver": np.array(solver), "steps": np.array(steps)}
    log = io.StringIO()
    t0 = time.time()
    cwd = os.getcwd()
    os.chdir(os.path.dirname(os.path.dirname(HERE)))          # "./obj/cube1.STL" (ParticleSystem.py:42 loads it relative to cwd)
    try:
        with contextlib.redirect_stdout(log):
            ps = ps_mod.ParticleSystem(cfg)
            sol = getattr(sol_mod, solver + "_solver")(ps, cfg)
            rig = rig_mod.rigid_solver(ps, cfg) if rig_mod else None
    finally:
        os.chdir(cwd)
    n = ps.particle_num
    out["particle_num"] = np.array(n)
    out["boundary_particles_num"] = np.array(ps.boundary_particles_num)
    out["grid_num"] = np.array(ps.grid_num.to_list(), dtype=np.int32)
    out["boundary_pos"] = ps.boundary_particles.pos.to_numpy()
    out["boundary_volume"] = ps.boundary_particles.volume.to_numpy()
    out["lattice_pos"] = ps.fluid_particles.pos.to_numpy()
    if solver == "pcisph":
        out["pc_delta"] = np.array(sol.delta[None], dtype=np.float32)
        out["pc_beta"] = np.array(sol.beta, dtype=np.float64)
    if pert is not None:
        pos0, vel = perturbed_state(out["lattice_pos"], *pert, diameter=2 * cfg["scene"]["particle_radius"])
        ps.fluid_particles.pos.from_numpy(pos0)
        ps.fluid_particles.vel.from_numpy(vel)
    out["pos0"] = ps.fluid_particles.pos.to_numpy()
    out["vel0"] = ps.fluid_particles.vel.to_numpy()

    def rigid_state(tag):
        rp = ps.rigid_particles
        for f in ("pos", "vel", "acc", "force", "omega", "alpha", "volume", "mass"):
            out["rigid_%s_%s" % (f, tag)] = getattr(rp, f).to_numpy()
        out["rigid_centroid_%s" % tag] = ps.rigid_centriod.to_numpy().reshape(3)
        out["rigid_inertia_inv_%s" % tag] = ps.rigid_inertia_tensor_inv.to_numpy().reshape(9)
        out["rigid_vertices_%s" % tag] = ps.rigid_vertices.to_numpy()
    if rig is not None:
        out["rigid_inertia"] = ps.rigid_inertia_tensor.to_numpy().reshape(9)
        if rigid_vel is not None:
            ps.rigid_particles.vel.fill(ti.Vector(rigid_vel))
        rigid_state("0")
    for s in range(1, steps + 1):
        log.seek(0)
        log.truncate()
        with contextlib.redirect_stdout(log):
            sol.step()
            if rig is not None:                       # main.py:166-171: the fluid sub-steps, then the rigid ones
                out["rigid_force_fluid_%d" % s] = ps.rigid_particles.force.to_numpy()
                rig.step()
                rigid_state(str(s))
                out["rs_omega_%d" % s] = rig.omega.to_numpy().reshape(3)
                out["rs_attitude_%d" % s] = rig.attitude.to_numpy().reshape(3)
                out["rs_dt_%d" % s] = np.array(rig.delta_time[None], dtype=np.float32)
        text = log.getvalue()
        out["pos_%d" % s] = ps.fluid_particles.pos.to_numpy()
        out["vel_%d" % s] = ps.fluid_particles.vel.to_numpy()
        out["cell3_%d" % s] = ps.fluid_particles.belong_grid.to_numpy()
        for f in FIELDS[solver]:
            out["%s_%d" % (f, s)] = getattr(sol, f).to_numpy()
        out["delta_time_%d" % s] = np.array(sol.delta_time[None], dtype=np.float32)
        for key, pat in LOG_PATTERNS.items():
            m = pat.findall(text)
            if m:
                out["log_%s_%d" % (key, s)] = np.array([float(x) for x in m[-1]], dtype=np.float64)
        if name.endswith("_visualize") and s == steps:
            ps.reset_grid()
            ps.update_grid()
            sol.visualize_rho()
            out["rgb_rho"] = ps.rgb.to_numpy()
            sol.visualize_neighbour()
            out["rgb_neighbour"] = ps.rgb.to_numpy()
        sys.stderr.write("%s step %d: %.0f s  %s\n" % (name, s, time.time() - t0, " | ".join(
            l.strip() for l in text.splitlines() if "iteration" in l or "Iter cnt" in l)))
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    out_path = None
    if "--out" in sys.argv:
        out_path = sys.argv[sys.argv.index("--out") + 1]
        args = [a for a in args if a != out_path]
    for name in args or list(CASES):
        res = run_case(name)
        path = out_path or os.path.join(HERE, "refshim_%s.npz" % name)
        np.savez_compressed(path, **res)
        sys.stderr.write("wrote %s (%d bytes)\n" % (path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
