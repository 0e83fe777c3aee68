"""Parity at the sizes BASELINE.json names: configs[1] DFSPH 10^6 particles, configs[2] PCISPH /
IISPH 4.1 M particles, configs[3] the 1000-step rigid-pose statistic.

The strict kernels must be BIT-EXACT against the oracle at these sizes too, with identical iteration counts --
which also answers the reduction-order question: the device-side loop averages are accumulated in float64 (block
partials + tree), the oracle's in the reference's one-thread float32 order; they differ by ~1e-4 relative at 10^6
particles, and the decisions (thresholds 10 for the divergence average DF:400, 0.1 for the density error DF:225)
must -- and do -- come out identical.  The margin is asserted below.  The oracle runs with OpenMP over particles
(per-particle arithmetic is order-preserving; only its loop averages are a serial float32 sum)."""
import os

import numpy as np
import pytest

from cfd_taichi_b200 import scenes, selfcheck
from cfd_taichi_b200.dfsph_solver import dfsph_solver
from cfd_taichi_b200.iisph_solver import iisph_solver
from cfd_taichi_b200.pcisph_solver import pcisph_solver
from cfd_taichi_b200.rigid_solver import rigid_solver
from conftest import ROOT, quiet_ps, quiet_solver
from oracle import oracle as O

pytestmark = pytest.mark.gpu
THREADS = len(os.sched_getaffinity(0))


def test_dfsph_one_million_strict_bit_exact_three_steps(built):
    """configs[1]: 100^3 particles, 3 steps (the third runs the full 15-pass divergence loop)."""
    cfg = scenes.breaking_dam(100)
    ps = quiet_ps(cfg, strict=True, solver_name="dfsph")
    sol = quiet_solver(dfsph_solver, ps, cfg)
    o = O.Oracle(cfg, solver="dfsph", threads=THREADS)
    assert ps.particle_num == 1000000
    seen = []
    for step in range(3):
        sol.step(); o.step()
        st = sol.stats()
        it_o = (int(o.scalar("df_div_iters")), int(o.scalar("df_den_iters")))
        assert st.error_flags == 0
        assert (st.div_iters, st.den_iters) == it_o, "step %d: GPU %s, oracle %s" % (step, (st.div_iters, st.den_iters), it_o)
        seen.append(it_o)
        assert st.delta_time == np.float32(o.scalar("delta_time"))
        for nm, a, b in [("pos", ps.fluid_particles.pos.to_numpy(), o.field("pos")), ("vel", ps.fluid_particles.vel.to_numpy(), o.field("vel")),
                         ("rho", sol.rho.to_numpy(), o.field("rho")), ("alpha", sol.alpha.to_numpy(), o.field("alpha")),
                         ("rho_derivative", sol.rho_derivative.to_numpy(), o.field("rho_derivative")),
                         ("rho_adv", sol.rho_adv.to_numpy(), o.field("rho_adv")),
                         ("warm_start_k", sol.warm_start_k.to_numpy(), o.field("warm_start_k"))]:
            assert np.array_equal(a, b), "step %d: %s differs at 10^6 particles (rel %.3e)" % (step, nm, selfcheck.relinf(a, b))
        assert np.array_equal(ps.neighbour_counts().cpu().numpy(), o.field("nbr_count"))
        # float64 tree (GPU) against one-thread float32 (oracle) averages: close, and far from the thresholds
        d_gpu, d_orc = float(st.div_err), float(o.scalar("df_div_err"))
        assert abs(d_gpu - d_orc) <= 2e-3 * max(1.0, abs(d_orc)), (d_gpu, d_orc)
        if it_o[0] == 15:
            assert abs(d_orc - 10.0) > 100 * abs(d_gpu - d_orc), "divergence average %.6g too close to DF:400's threshold" % d_orc
        e_gpu, e_orc = float(st.den_err), float(o.scalar("df_den_err"))
        assert abs(e_gpu - e_orc) <= 1e-3, (e_gpu, e_orc)
        assert abs(e_orc - 0.1) > 20 * abs(e_gpu - e_orc), "density error %.6g too close to DF:225's threshold" % e_orc
    assert any(s[0] == 15 for s in seen)
    ps.close(); o.close()


@pytest.mark.parametrize("solver,cls,dt", [("pcisph", pcisph_solver, 1.5e-4), ("iisph", iisph_solver, 2.5e-4)])
def test_four_million_strict_bit_exact(built, solver, cls, dt):
    """configs[2]: 160^3 = 4 096 000 particles, 2 steps of the iterative pressure solve."""
    cfg = scenes.breaking_dam(160, solver, dt)
    ps = quiet_ps(cfg, strict=True, solver_name=solver)
    sol = quiet_solver(cls, ps, cfg)
    o = O.Oracle(cfg, solver=solver, threads=THREADS)
    assert ps.particle_num == 4096000 and ps.boundary_particles_num == 465122
    if solver == "pcisph":
        assert sol.delta[None] == np.float32(o.scalar("pc_delta"))
    for step in range(2):
        sol.step(); o.step()
        st = sol.stats()
        assert st.error_flags == 0
        it_gpu = st.pc_iters if solver == "pcisph" else st.ii_iters
        it_orc = int(o.scalar("pc_iters" if solver == "pcisph" else "ii_iters"))
        assert it_gpu == it_orc, "step %d: %d iterations on the GPU, %d in the oracle" % (step, it_gpu, it_orc)
        for nm, a, b in [("pos", ps.fluid_particles.pos.to_numpy(), o.field("pos")), ("vel", ps.fluid_particles.vel.to_numpy(), o.field("vel")),
                         ("rho", sol.rho.to_numpy(), o.field("rho"))]:
            assert np.array_equal(a, b), "step %d: %s differs at 4 M particles (rel %.3e)" % (step, nm, selfcheck.relinf(a, b))
        press = (sol.press_iter.to_numpy(), o.field("press_iter")) if solver == "pcisph" else (sol.p_iter.to_numpy(), o.field("p_iter"))
        assert np.array_equal(*press)
    ps.close(); o.close()


def _rigid_run(strict, steps, perturb=False):
    """dam_flush_cube with DFSPH (configs[3]) for `steps` steps; per step (t, centroid xyz, body velocity xyz, fluid
    kinetic energy, mean density), t = accumulated adaptive delta_time (DF:112-119)."""
    cfg = scenes.shipped("dam_flush_cube", "dfsph")
    ps = quiet_ps(cfg, strict=strict, solver_name="dfsph", base_dir=ROOT)
    sol = quiet_solver(dfsph_solver, ps, cfg)
    rs = rigid_solver(ps, cfg)
    n = ps.particle_num
    if perturb:
        x = ps._pos4[4321, 1].item()
        ps._pos4[4321, 1] = float(np.nextafter(np.float32(x), np.float32(10)))
    rows, t = [], 0.0
    for k in range(steps):
        sol.step(); rs.step()
        st = sol.stats()
        assert st.error_flags == 0, "step %d flags %d" % (k, st.error_flags)
        t += st.delta_time
        info = ps.rigid_state()
        v = ps._vel4[:n, :3].double()
        rows.append([t] + list(info.centroid) + list(info.vel) + [0.5 * 0.125 * (v * v).sum().item(),
                                                                  sol.rho.to_torch().double().mean().item()])
    ps.close()
    return np.array(rows)


def test_rigid_pose_after_1000_steps(built):
    """BASELINE.json: after 1000 steps the rigid-body pose, kinetic energy and mean density agree within 1 %.
    The strict kernels stand in for the oracle (bit-exact against it on this scene: tests/test_gpu_rigid.py,
    test_gpu_fast_parity.py; 1000 oracle steps of 56 k particles take minutes).  DFSPH's time step adapts to the
    fastest particle, so two runs drift apart in simulated time per step: the statistics are compared at equal
    SIMULATED TIME (the end of the shortest run).  The chaos floor of the reference's own arithmetic is measured
    alongside (one coordinate of one particle moved by one ulp in the strict run)."""
    steps = 1000
    ref = _rigid_run(True, steps)
    ulp = _rigid_run(True, steps, perturb=True)
    fast = _rigid_run(False, steps)
    T = min(ref[-1, 0], ulp[-1, 0], fast[-1, 0])
    box = np.array([5.0, 3.0, 1.5])

    def at(run, col):
        return float(np.interp(T, run[:, 0], run[:, col]))

    def report(tag, run):
        dc = max(abs(at(run, 1 + k) - at(ref, 1 + k)) / box[k] for k in range(3))       # pose, relative to the box
        vmax = np.abs(ref[:, 4:7]).max() + 1e-30
        dv = max(abs(at(run, 4 + k) - at(ref, 4 + k)) for k in range(3)) / vmax
        dke = abs(at(run, 7) - at(ref, 7)) / at(ref, 7)
        drho = abs(at(run, 8) - at(ref, 8)) / at(ref, 8)
        print("%s at t = %.4f s (%d steps): centroid %.3e of the box, body velocity %.3e, kinetic energy %.3e, "
              "mean density %.3e" % (tag, T, steps, dc, dv, dke, drho))
        return dc, dv, dke, drho

    print()
    floor = report("strict, one ulp moved ", ulp)
    got = report("fast kernels          ", fast)
    print("the body's centroid moved %.4f m in the reference run" % np.abs(ref[-1, 1:4] - ref[0, 1:4]).max())
    assert got[0] <= 0.01 and got[3] <= 0.01          # pose and mean density within 1 %
    assert got[2] <= max(0.01, 3.0 * floor[2])          # kinetic energy within 1 % or the reference's own chaos floor
    assert got[1] <= max(0.01, 3.0 * floor[1])
