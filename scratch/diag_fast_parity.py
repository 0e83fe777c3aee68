"""One-substep error of the fast (and strict) kernels against the oracle, from states the strict run reached.
Diagnostic only (prints a table); the asserted version lives in tests/test_gpu_fast_parity.py."""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cfd_taichi_b200 import scene, scenes  # noqa: E402
from cfd_taichi_b200 import main as app  # noqa: E402
from cfd_taichi_b200.ParticleSystem import ParticleSystem  # noqa: E402
from cfd_taichi_b200.dfsph_solver import dfsph_solver  # noqa: E402
from cfd_taichi_b200.iisph_solver import iisph_solver  # noqa: E402
from cfd_taichi_b200.pcisph_solver import pcisph_solver  # noqa: E402
from cfd_taichi_b200.wcsph_solver import wcsph_solver  # noqa: E402
from cfd_taichi_b200.rigid_solver import rigid_solver  # noqa: E402
from oracle import oracle as O  # noqa: E402

CLS = {"dfsph": dfsph_solver, "wcsph": wcsph_solver, "pcisph": pcisph_solver, "iisph": iisph_solver}


def relinf(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / (np.abs(b).max() + 1e-30))


def box_points(lo, hi, pitch=0.05):
    ax = [np.arange(int(round(lo[k] / pitch)), int(round(hi[k] / pitch)) + 1) * pitch for k in range(3)]
    return np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


FIELDS = {
    "dfsph": [("rho", "rho"), ("alpha", "alpha"), ("rho_derivative", "rho_derivative"), ("rho_adv", "rho_adv"),
              ("vel_adv", "vel_adv"), ("force_ext", "force_ext"), ("warm_start_k", "warm_start_k")],
    "wcsph": [("rho", "rho"), ("pressure", "pressure"), ("pressure_gradient", "pressure_gradient"),
              ("viscosity", "viscosity"), ("tension", "tension"), ("boundary_acc", "boundary_acc")],
    "pcisph": [("rho", "rho"), ("ext_force", "ext_force"), ("press_force", "press_force"), ("press_iter", "press_iter"),
               ("rho_err", "rho_err"), ("pos_predict", "pos_predict")],
    "iisph": [("rho", "rho"), ("f_adv", "f_adv"), ("v_adv", "v_adv"), ("d_ii", "d_ii"), ("a_ii", "a_ii"),
              ("rho_adv", "rho_adv"), ("p_iter", "p_iter"), ("d_ij", "d_ij"), ("r_sum", "r_sum"), ("f_press", "f_press")],
}


def iters(solver, st, o):
    if solver == "dfsph":
        return (st.div_iters, st.den_iters), (int(o.scalar("df_div_iters")), int(o.scalar("df_den_iters")))
    if solver == "pcisph":
        return st.pc_iters, int(o.scalar("pc_iters"))
    if solver == "iisph":
        return st.ii_iters, int(o.scalar("ii_iters"))
    return 0, 0


def run(scene_name, solver, warms, rigid=False):
    cfg = scenes.shipped(scene_name, solver)
    pts = verts = None
    if rigid:
        pts = box_points([0, 0, 0], [0.8, 0.5, 1.0])
        verts = np.array([[x, y, z] for x in (0, 0.8) for y in (0, 0.5) for z in (0, 1.0)], dtype=np.float32)
        scene.rigid_points_from_config = lambda solid, base_dir=".": (pts, verts, None)
    else:
        cfg.pop("solid", None)
    mk = lambda strict: quiet(ParticleSystem, cfg, strict=strict, solver_name=solver)
    ps_s, ps_f = mk(True), mk(False)
    sol_s, sol_f = quiet(CLS[solver], ps_s, cfg), quiet(CLS[solver], ps_f, cfg)
    rs_s = rigid_solver(ps_s, cfg) if rigid else None
    rs_f = rigid_solver(ps_f, cfg) if rigid else None
    o = O.Oracle(cfg, solver=solver, rigid_points=pts, rigid_vertices=verts, threads=1 if rigid else 8)
    done = 0
    tmp = tempfile.mkdtemp()
    for warm in warms:
        while done < warm:
            sol_s.step()
            if rs_s:
                rs_s.step()
            o.step()
            done += 1
        assert np.array_equal(ps_s.fluid_particles.pos.to_numpy(), o.field("pos")), "strict run left the oracle"
        app.save_state(os.path.join(tmp, "s"), ps_s, sol_s, rs_s)
        app.load_state(os.path.join(tmp, "s"), ps_f, sol_f)
        sol_f.step(); sol_s.step()
        import ctypes
        O.lib().orc_step(o._h)
        frc = (relinf(ps_f.rigid_particles.force.to_numpy(), o.field("rforce")),) if rigid else ()
        if rs_s:
            rs_s.step(); rs_f.step()
        O.lib().orc_rigid_step(o._h)
        done += 1
        st = sol_f.stats()
        a, b = iters(solver, st, o)
        out = ["%s/%s%s warm=%d flags=%d iters fast=%s oracle=%s" % (scene_name, solver, "+rigid" if rigid else "", warm,
                                                                    st.error_flags, a, b)]
        worst = 0.0
        for nm, on in FIELDS[solver]:
            e = relinf(getattr(sol_f, nm).to_numpy(), o.field(on))
            es = relinf(getattr(sol_s, nm).to_numpy(), o.field(on))
            out.append("   %-18s fast %.2e   strict %.2e   max|ref| %.3e" % (nm, e, es, np.abs(o.field(on)).max()))
            worst = max(worst, e)
        for nm, x, y, xs in [("pos", ps_f.fluid_particles.pos.to_numpy(), o.field("pos"), ps_s.fluid_particles.pos.to_numpy()),
                             ("vel", ps_f.fluid_particles.vel.to_numpy(), o.field("vel"), ps_s.fluid_particles.vel.to_numpy())]:
            e = relinf(x, y)
            out.append("   %-18s fast %.2e   strict %.2e   max|ref| %.3e" % (nm, e, relinf(xs, y), np.abs(y).max()))
            worst = max(worst, e)
        if rigid:
            out.append("   %-18s fast %.2e" % ("rigid force", frc[0]))
            cf, co = np.array(list(ps_f.rigid_state().centroid)), o.field("centroid").reshape(-1)
            out.append("   %-18s fast %.2e" % ("rigid centroid", relinf(cf, co)))
            out.append("   %-18s fast %.2e" % ("rigid rpos", relinf(ps_f.rigid_particles.pos.to_numpy(), o.field("rpos"))))
        out[0] += "  WORST %.2e %s" % (worst, "OK" if worst <= 1e-5 and a == b else "MISS")
        print("\n".join(out), flush=True)
    ps_s.close(); ps_f.close(); o.close()


if __name__ == "__main__":
    run("small_block", "dfsph", [0, 1, 20, 100, 300])
    run("breaking_dam_30k", "dfsph", [0, 5, 40])
    run("dam_flush_cube", "dfsph", [0, 3], rigid=True)
    for s in ("wcsph", "pcisph", "iisph"):
        run("small_block", s, [0, 1, 20, 100])
    run("breaking_dam_30k", "wcsph", [0, 20])
    run("breaking_dam_30k", "iisph", [0, 10])
    run("breaking_dam_30k", "pcisph", [0, 10])
