"""Generates reference-pinned fixtures: one-substep fields of the UNMODIFIED reference (Jukgei/CFD_Taichi) under
ti.cpu, dumped as tests/golden/reference_<scene>_<solver>.npz.

    python tests/golden/make_reference_golden.py [--reference /root/reference] [--steps 3]

STATUS: cannot run in the build image -- `taichi==1.6.0` (reference requirements.txt:1) has no CPython 3.12 wheel
in the offline wheelhouse and there is no network, so no fixture produced by this script is committed yet and the
oracle stays "parity unpinned" (oracle/sph_oracle.h).  It is committed because it is the one
route from "partial" to pinned parity: on any machine with taichi 1.6 (CPython <= 3.11) and the reference checkout
it writes the fixtures, and tests/test_golden.py::test_oracle_against_reference_fixtures then holds the oracle to
them -- integers (cell ids, sorted order, neighbour counts, iteration counts) bit-exactly, floats at 1e-6 relative
(the LLVM CPU backend may contract a * b + c) -- with no further change.

What it does, per (scene file, solver): ti.init(arch=ti.cpu, cpu_max_num_threads=1, fast_math=False,
default_fp=ti.f32) -- one thread makes the atomic accumulations of for_all_neighbor / the averages ordered, i.e.
the canonical order the oracle restates -- builds ParticleSystem + <name>_solver exactly like main.py:64-71,
runs `steps` times solver.step() (+ rigid_solver.step()) and after every step records the particle state, the
solver's per-particle fields, the grid arrays and the loop statistics the solver prints.
"""
import argparse
import contextlib
import importlib
import io
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# (fixture name, config file relative to the reference root, solver.name override or None)
CASES = [
    ("small_block_dfsph", "config/dfsph_config_backup.json", None),
    ("small_block_wcsph", "config/wcsph_config_backup.json", None),
    ("small_block_pcisph", "config/pcisph_config_backup.json", None),
    ("small_block_iisph", "config/iisph_config_backup.json", None),
    ("breaking_dam_30k_wcsph", "config/breaking_dam_30k.json", "wcsph"),      # BASELINE.json configs[0]
    ("breaking_dam_30k_dfsph", "config/breaking_dam_30k.json", "dfsph"),
    ("dam_flush_cube_dfsph", "config/dam_flush_cube.json", "dfsph"),          # BASELINE.json configs[3]
]
# per-particle solver fields worth pinning (attribute names of the reference classes)
SOLVER_FIELDS = {
    "dfsph": ["rho", "alpha", "rho_adv", "rho_derivative", "vel_adv", "force_ext", "warm_start_k"],
    "wcsph": ["rho", "pressure", "pressure_gradient", "viscosity", "tension", "boundary_acc"],
    "pcisph": ["rho", "ext_force", "press_force", "press_iter", "rho_err", "pos_predict", "vel_predict"],
    "iisph": ["rho", "f_adv", "v_adv", "d_ii", "a_ii", "rho_adv", "p_iter", "p_past", "d_ij", "r_sum", "f_press"],
}
# the loop statistics are only printed by the reference (DF:416, DF:233, PC:70, II:100): parse its stdout
STAT_PATTERNS = {
    "div_iters": re.compile(r"\[divergence iteration\] count: (\d+)"),
    "div_first_err": re.compile(r"\[divergence iteration\] count: \d+, first error ([-\w.+]+)"),
    "div_err": re.compile(r"\[divergence iteration\] count: \d+, first error [-\w.+]+, error ([-\w.+]+)"),
    "den_iters": re.compile(r"\[density iteration\] count: (\d+)"),
    "den_err": re.compile(r"\[density iteration\] count: \d+, error ([-\w.+]+)"),
    "iter_cnt": re.compile(r"Iter cnt: +(\d+)"),
}


def to_np(x):
    return x.to_numpy() if hasattr(x, "to_numpy") else np.asarray(x)


def run_case(ref_root, name, config_rel, solver_override, steps, out_dir):
    import taichi as ti
    ti.init(arch=ti.cpu, cpu_max_num_threads=1, fast_math=False, default_fp=ti.f32, debug=False)
    import utils                                       # the reference's own modules (sys.path[0] = ref_root)
    from ParticleSystem import ParticleSystem
    from rigid_solver import rigid_solver
    config = utils.read_config(os.path.join(ref_root, config_rel))
    if solver_override:
        config["solver"]["name"] = solver_override
    solver_name = config["solver"]["name"]
    cwd = os.getcwd()
    os.chdir(ref_root)                                 # mesh paths in the scene files are relative to the checkout
    try:
        ps = ParticleSystem(config)
        module = importlib.import_module(solver_name + "_solver")           # main.py:65-68
        solver = getattr(module, solver_name + "_solver")(ps, config)
        rs = rigid_solver(ps, config) if config.get("solid", {}) else None   # main.py:70-71
        out = {"config_json": np.frombuffer(json.dumps(config, sort_keys=True).encode(), dtype=np.uint8),
               "particle_num": np.int64(ps.particle_num), "boundary_particles_num": np.int64(ps.boundary_particles_num),
               "grid_num": np.asarray(list(ps.grid_num), dtype=np.int64),
               "boundary_pos": to_np(ps.boundary_particles.pos), "boundary_volume": to_np(ps.boundary_particles.volume),
               "pos_0": to_np(ps.fluid_particles.pos), "vel_0": to_np(ps.fluid_particles.vel)}
        if rs is not None:
            out["rigid_pos_0"] = to_np(ps.rigid_particles.pos)
            out["rigid_volume"] = to_np(ps.rigid_particles.volume)
            out["rigid_mass"] = to_np(ps.rigid_particles.mass)
            out["rigid_centroid_0"] = to_np(ps.rigid_centriod)
            out["rigid_inertia_tensor"] = to_np(ps.rigid_inertia_tensor)
        for k in range(1, steps + 1):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                solver.step()                                               # main.py:166-167
            # fluid -> rigid forces as accumulated by the fluid step, before the rigid step consumes them
            if rs is not None:
                out["rigid_force_%d" % k] = to_np(ps.rigid_particles.force)
                if ps.active_rigid[None] == 1:
                    rs.step()                                               # main.py:169-171
            text = buf.getvalue()
            for key, pat in STAT_PATTERNS.items():
                m = pat.search(text)
                if m:
                    out["%s_%d" % (key, k)] = np.float64(float(m.group(1)))
            out["delta_time_%d" % k] = np.float32(solver.delta_time[None])
            out["pos_%d" % k] = to_np(ps.fluid_particles.pos)
            out["vel_%d" % k] = to_np(ps.fluid_particles.vel)
            out["belong_grid_%d" % k] = to_np(ps.fluid_particles.belong_grid)
            for f in SOLVER_FIELDS.get(solver_name, []):
                if hasattr(solver, f):
                    out["%s_%d" % (f, k)] = to_np(getattr(solver, f))
            if rs is not None:
                out["rigid_pos_%d" % k] = to_np(ps.rigid_particles.pos)
                out["rigid_centroid_%d" % k] = to_np(ps.rigid_centriod)
        # neighbour counts of the LAST step's start positions are what solver fields were computed from; the
        # count kernel is a @ti.func, so evaluate it through the public arg-max helper's building block
        if hasattr(ps, "get_neighbour_count"):
            @ti.kernel
            def counts(dst: ti.types.ndarray()):
                for i in range(ps.particle_num):
                    dst[i] = ps.get_neighbour_count(i)
            c = np.zeros(ps.particle_num, dtype=np.int32)
            ps.reset_grid(); ps.update_grid()
            counts(c)
            out["neighbour_count_after_%d" % steps] = c
    finally:
        os.chdir(cwd)
    path = os.path.join(out_dir, "reference_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("wrote", path, "(%d arrays)" % len(out))
    ti.reset()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--only", default=None, help="fixture name (default: all)")
    ap.add_argument("--out", default=HERE)
    args = ap.parse_args()
    try:
        import taichi  # noqa: F401
    except ImportError:
        raise SystemExit("make_reference_golden.py: taichi is not importable here (the reference pins taichi==1.6.0, "
                         "no CPython 3.12 wheel offline); run this on a machine that has it")
    if not os.path.isdir(args.reference):
        raise SystemExit("reference checkout not found: " + args.reference)
    sys.path.insert(0, args.reference)
    for name, cfg, solver in CASES:
        if args.only and args.only != name:
            continue
        run_case(args.reference, name, cfg, solver, args.steps, args.out)


if __name__ == "__main__":
    main()
