"""Sweep-by-sweep comparison of the two arithmetic modes of the CUDA library (strict IEEE / fast FMA).

The reference's divergence-free loop (dfsph_solver.py:393-416) is not a contraction on a collapsing dam:
it always runs to its cap of 15 passes and multiplies a 1-ulp input difference by ~1e4 (measured by
`step_sensitivity`, asserted in tests/test_gpu_fast_parity.py).  A whole-substep comparison of two
implementations that are not bit-identical therefore measures the conditioning of the reference's own
loop, not the kernels.  What CAN be held to BASELINE.json's 1e-5 is every sweep in isolation: each kernel
of the fast path is fed the strict path's exact inputs (`sph_copy_work_state`) and its outputs are compared.
The strict path itself is bit-exact against the oracle (tests), so strict inputs == oracle inputs.

Nothing here touches oracle/; tests and bench.py's `parity` block add the oracle comparison on top.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def relinf(a, b):
    """||a - b||_inf / ||b||_inf (BASELINE.json: "within 1e-5 relative")."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    same = (a == b) | (np.isnan(a) & np.isnan(b))      # a work array no sweep has written yet is equal garbage
    if same.all():
        return 0.0
    with np.errstate(invalid="ignore"):
        d = np.where(same, 0.0, np.abs(a - b))
    if not np.isfinite(d).all():
        return float("inf")
    return float(d.max() / (np.abs(b[np.isfinite(b)]).max() + 1e-30))


def copy_work_state(dst_ps, src_ps):
    _lib.check(dst_ps._lib.sph_copy_work_state(dst_ps._h, src_ps._h, dst_ps._stream()), dst_ps._h)


def copy_caller_state(dst_ps, dst_sol, src_ps, src_sol):
    """Everything a step depends on (what main.save_state dumps), device to device: positions, velocities with
    the solver-persistent scalar in .w, the time step, and the rigid body (particles, vertices, body state)."""
    dst_ps._pos4.copy_(src_ps._pos4)
    dst_ps._vel4.copy_(src_ps._vel4)
    dst_sol.delta_time[None] = src_sol.delta_time[None]
    if src_ps.exist_rigid[None]:
        for name in ("_rpos4", "_rvel4", "_rforce4", "_rverts4"):
            getattr(dst_ps, name).copy_(getattr(src_ps, name))
        info = src_ps.rigid_state()
        _lib.check(dst_ps._lib.sph_rigid_set_state(dst_ps._h, ctypes.byref(info)), dst_ps._h)


def _f(ps, fid, width=1):
    return ps._fetch(fid, width, torch.float32).cpu().numpy()


def _cmp(out, piece, pairs):
    rec = out.setdefault(piece, {})
    for name, a, b in pairs:
        rec[name] = max(rec.get(name, 0.0), relinf(a, b))


def dfsph_sweeps(ps_s, sol_s, ps_f, sol_f, rigid=False, max_den=64):
    """One DFSPH step on a strict and a fast handle that hold the SAME caller state, executed sweep by
    sweep (DF:423-438 with the two loops unrolled one pass at a time); before every sweep the fast handle
    receives the strict handle's work state.  Returns (errors, info): errors[piece][field] = relinf of the
    fast output against the strict output (max over the passes of a loop), info = iteration counts and
    loop residuals of both modes."""
    err, info = {}, {}
    for ps, sol in ((ps_s, sol_s), (ps_f, sol_f)):
        sol.simulate_cnt[None] += 1
        ps.update_grid()
        sol.initialize()
    n_s, n_f = ps_s.neighbour_counts().cpu().numpy(), ps_f.neighbour_counts().cpu().numpy()
    info["neighbour_counts_equal"] = bool(np.array_equal(n_s, n_f))
    _cmp(err, "initialize", [("rho", _f(ps_f, _lib.F_RHO), _f(ps_s, _lib.F_RHO)),
                             ("alpha", _f(ps_f, _lib.F_ALPHA), _f(ps_s, _lib.F_ALPHA))])

    def both(phase):
        copy_work_state(ps_f, ps_s)
        ps_s.phase(phase)
        ps_f.phase(phase)
        return ps_s.read_stats(), ps_f.read_stats()

    def cmp_div(piece):
        vs, vf = _f(ps_s, _lib.F_FLUID_VEL, 4), _f(ps_f, _lib.F_FLUID_VEL, 4)
        _cmp(err, piece, [("vel", vf[:, :3], vs[:, :3]), ("warm_start_k", vf[:, 3], vs[:, 3]),
                          ("rho_derivative", _f(ps_f, _lib.F_RHO_DERIVATIVE), _f(ps_s, _lib.F_RHO_DERIVATIVE))])

    def drho_scale():   # the averages are compared on the scale of the values they average
        return float(np.abs(_f(ps_s, _lib.F_RHO_DERIVATIVE)).max())

    ss, sf = both(_lib.PH_DF_DIV_BEGIN)
    cmp_div("divergence_warm_start+derivative")
    _cmp(err, "divergence_warm_start+derivative", [("avg", [sf.div_first_err, drho_scale()], [ss.div_first_err, drho_scale()])])
    flags_equal = ss.div_active == sf.div_active
    passes = 0
    while ss.div_active and passes < 15:
        ss, sf = both(_lib.PH_DF_DIV_ONE)
        passes += 1
        cmp_div("divergence_iter+derivative")
        _cmp(err, "divergence_iter+derivative", [("avg", [sf.div_err, drho_scale()], [ss.div_err, drho_scale()])])
        flags_equal = flags_equal and ss.div_active == sf.div_active and ss.div_iters == sf.div_iters
    info["div_iters"] = (ss.div_iters, sf.div_iters)

    ss, sf = both(_lib.PH_DF_EXT_FORCE_VEL_ADV)
    _cmp(err, "ext_force+vel_adv", [("force_ext", _f(ps_f, _lib.F_FORCE_A, 4)[:, :3], _f(ps_s, _lib.F_FORCE_A, 4)[:, :3]),
                                    ("vel_adv", _f(ps_f, _lib.F_VEL_ADV, 4)[:, :3], _f(ps_s, _lib.F_VEL_ADV, 4)[:, :3]),
                                    ("delta_time", [sf.delta_time], [ss.delta_time])])
    passes = 0
    while True:
        ss, sf = both(_lib.PH_DF_DEN_ONE)
        passes += 1
        pairs = [("rho_adv", _f(ps_f, _lib.F_RHO_ADV), _f(ps_s, _lib.F_RHO_ADV)),
                 ("vel_adv", _f(ps_f, _lib.F_VEL_ADV, 4)[:, :3], _f(ps_s, _lib.F_VEL_ADV, 4)[:, :3]),
                 ("avg", [sf.den_err + 1000.0], [ss.den_err + 1000.0])]    # DF:225 compares the average density
        if rigid:
            pairs.append(("rigid_force", ps_f.rigid_particles.force.to_numpy(), ps_s.rigid_particles.force.to_numpy()))
        _cmp(err, "rho_adv+iter_vel_adv", pairs)
        flags_equal = flags_equal and ss.den_active == sf.den_active and ss.den_iters == sf.den_iters
        if not ss.den_active or passes >= max_den:
            break
    info["den_iters"] = (ss.den_iters, sf.den_iters)
    info["loop_flags_equal"] = bool(flags_equal)

    both(_lib.PH_DF_POSITION)
    _cmp(err, "position", [("pos", ps_f.fluid_particles.pos.to_numpy(), ps_s.fluid_particles.pos.to_numpy()),
                           ("vel", ps_f.fluid_particles.vel.to_numpy(), ps_s.fluid_particles.vel.to_numpy())])
    info["error_flags"] = (ps_s.read_stats().error_flags, ps_f.read_stats().error_flags)
    return err, info


def _fields(ps, spec):
    out = {}
    for name, fid, width in spec:
        a = _f(ps, fid, 4 if width == 3 else 1)
        out[name] = a[:, :3] if width == 3 else a.reshape(-1)
    return out


_PC_FIELDS = [("pos_predict", _lib.F_VEC_C, 3), ("vel_predict", _lib.F_VEL_ADV, 3), ("ext_force", _lib.F_FORCE_A, 3),
              ("press_force", _lib.F_FORCE_B, 3), ("rho_predict", _lib.F_SCALAR_A, 1), ("rho_err", _lib.F_SCALAR_B, 1),
              ("press_iter", _lib.F_PRESSURE, 1), ("rho", _lib.F_RHO, 1)]
_II_FIELDS = [("v_adv", _lib.F_VEL_ADV, 3), ("f_adv", _lib.F_FORCE_A, 3), ("d_ii", _lib.F_VEC_A, 3), ("a_ii", _lib.F_SCALAR_A, 1),
              ("d_ij", _lib.F_FORCE_B, 3), ("rho_adv", _lib.F_RHO_ADV, 1), ("p_iter", _lib.F_PRESSURE, 1),
              ("r_sum", _lib.F_SCALAR_B, 1), ("rho", _lib.F_RHO, 1)]
_WC_FIELDS = [("pressure", _lib.F_PRESSURE, 1), ("pressure_gradient", _lib.F_FORCE_A, 3), ("viscosity", _lib.F_FORCE_B, 3),
              ("tension", _lib.F_VEC_A, 3), ("boundary_acc", _lib.F_VEC_B, 3), ("rho", _lib.F_RHO, 1)]


def _generic_sweeps(ps_s, sol_s, ps_f, sol_f, first, fields, loop, last, rigid, iters_of, residual_of, max_passes):
    """first / last: phase ids; loop = (begin phase, one-pass phase) or None; see dfsph_sweeps for the contract."""
    err, info = {}, {}
    for ps, sol in ((ps_s, sol_s), (ps_f, sol_f)):
        sol.simulate_cnt[None] += 1
        ps.update_grid()

    def both(phase, piece, copy=True):
        if copy:
            copy_work_state(ps_f, ps_s)
        ps_s.phase(phase)
        ps_f.phase(phase)
        a, b = _fields(ps_f, fields), _fields(ps_s, fields)
        pairs = [(k, a[k], b[k]) for k in a]
        if rigid:
            pairs.append(("rigid_force", ps_f.rigid_particles.force.to_numpy(), ps_s.rigid_particles.force.to_numpy()))
        ss, sf = ps_s.read_stats(), ps_f.read_stats()
        if residual_of is not None:
            pairs.append(("residual", [residual_of(sf) + 1000.0], [residual_of(ss) + 1000.0]))   # average density error on the scale of rho_0
        _cmp(err, piece, pairs)
        return ss, sf

    both(first, "first_phase")      # builds each handle's own neighbour lists
    info["neighbour_counts_equal"] = bool(np.array_equal(ps_s.neighbour_counts().cpu().numpy(),
                                                         ps_f.neighbour_counts().cpu().numpy()))
    flags_equal = True
    if loop is not None:
        ss, sf = both(loop[0], "loop_begin")
        flags_equal = ss.loop_active == sf.loop_active
        passes = 0
        while ss.loop_active and passes < max_passes:
            ss, sf = both(loop[1], "loop_pass")
            passes += 1
            flags_equal = flags_equal and ss.loop_active == sf.loop_active and iters_of(ss) == iters_of(sf)
        info["iters"] = (iters_of(ss), iters_of(sf))
    info["loop_flags_equal"] = bool(flags_equal)
    copy_work_state(ps_f, ps_s)
    ps_s.phase(last)
    ps_f.phase(last)
    vs, vf = ps_s._vel4[:ps_s.particle_num].cpu().numpy(), ps_f._vel4[:ps_f.particle_num].cpu().numpy()
    _cmp(err, "integration", [("pos", ps_f.fluid_particles.pos.to_numpy(), ps_s.fluid_particles.pos.to_numpy()),
                              ("vel", vf[:, :3], vs[:, :3]), ("vel.w", vf[:, 3], vs[:, 3])])
    info["error_flags"] = (ps_s.read_stats().error_flags, ps_f.read_stats().error_flags)
    return err, info


def pcisph_sweeps(ps_s, sol_s, ps_f, sol_f, rigid=False):
    """PC:233-240 sweep by sweep: ext force | predict + predicted density | one pressure pass ... | integration."""
    return _generic_sweeps(ps_s, sol_s, ps_f, sol_f, _lib.PH_PC_EXT_FORCE, _PC_FIELDS,
                           (_lib.PH_PC_ITER_BEGIN, _lib.PH_PC_ITER_ONE), _lib.PH_PC_INTEGRATION, rigid,
                           lambda st: st.pc_iters, lambda st: st.pc_err, 80)


def iisph_sweeps(ps_s, sol_s, ps_f, sol_f, rigid=False):
    """II:342-349 sweep by sweep: predict_advection | one relaxed Jacobi pass ... | integration."""
    return _generic_sweeps(ps_s, sol_s, ps_f, sol_f, _lib.PH_II_PREDICT_ADVECTION, _II_FIELDS,
                           (_lib.PH_II_SOLVE_BEGIN, _lib.PH_II_SOLVE_ONE), _lib.PH_II_INTEGRATION, rigid,
                           lambda st: st.ii_iters, lambda st: st.ii_residual, 180)


def wcsph_sweeps(ps_s, sol_s, ps_f, sol_f, rigid=False):
    """WC:25-30: pressure phase | kinematic phase."""
    return _generic_sweeps(ps_s, sol_s, ps_f, sol_f, _lib.PH_WC_PRESSURE, _WC_FIELDS, None, _lib.PH_WC_KINEMATIC,
                           rigid, None, None, 0)


SWEEPS = {"dfsph": dfsph_sweeps, "pcisph": pcisph_sweeps, "iisph": iisph_sweeps, "wcsph": wcsph_sweeps}


def worst(err):
    """(value, 'piece/field') of the largest entry of a dfsph_sweeps error table."""
    best = (0.0, "")
    for piece, rec in err.items():
        for name, v in rec.items():
            if v >= best[0]:
                best = (v, piece + "/" + name)
    return best


def perturb_velocities_one_ulp(ps, seed=0):
    """Move every velocity component of the caller state by one ulp up or down (seeded)."""
    n = ps.particle_num
    v = ps._vel4[:n, :3].cpu().numpy()
    rng = np.random.default_rng(seed)
    up = rng.integers(0, 2, size=v.shape).astype(bool)
    w = np.where(up, np.nextafter(v, np.float32(np.inf)), np.nextafter(v, np.float32(-np.inf))).astype(np.float32)
    ps._vel4[:n, :3] = torch.from_numpy(w).to(ps._device)
