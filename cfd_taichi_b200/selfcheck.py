"""Sweep-by-sweep comparison of the two arithmetic modes of the CUDA library (strict IEEE / fast FMA).

Why sweep by sweep: the reference's solver loops are error amplifiers on a collapsing dam.  Its divergence-free
loop (dfsph_solver.py:393-416) always runs to its cap of 15 passes, and a difference operator follows every
velocity update (D rho / D t = sum m (v_i - v_j) . grad W cancels to a small residual), so a 1-ulp change of the
input velocities becomes ~1e-3 of the velocity scale after one substep -- with the STRICT kernels, i.e. in the
reference's own arithmetic (tests/test_gpu_fast_parity.py measures and asserts this).  A whole-substep comparison
of two implementations that are not bit-identical therefore measures the conditioning of the reference's loop, not
the kernels.  What can be held to BASELINE.json's 1e-5 is every sweep in isolation: each sweep of the fast path is
fed the strict path's exact inputs (`sph_copy_work_state`) and ALL work arrays, gathered payload copies included,
are compared afterwards.  The strict path is bit-exact against the oracle (tests), so strict inputs == oracle
inputs.

The loop statistics of the reference ("average over the particles whose value is > 0", DF:274-279, DF:141-149,
PC:123-133) jump when one particle crosses zero, so they are not compared between the modes; instead every
device-side average is checked against the float64 statistic of the SAME handle's per-particle field (that pins
the block-partial reduction), the per-particle fields are compared between the modes, and the loop decisions and
iteration counts must be identical.

Nothing here touches oracle/; tests and bench.py's `parity` block add the oracle comparison on top.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _t(x):
    """numpy array / list / torch tensor -> float64 torch tensor (device arrays stay on the device: at 1 M particles
    a sweep-by-sweep walk compares ~3 GB of work arrays, which only the GPU does in seconds)"""
    if isinstance(x, torch.Tensor):
        return x.to(torch.float64)
    return torch.as_tensor(np.asarray(x, dtype=np.float64))


def relinf(a, b, scale=None):
    """||a - b||_inf / ||b||_inf (BASELINE.json: "within 1e-5 relative"); `scale` replaces the denominator."""
    a, b = _t(a), _t(b)
    if a.numel() == 0:
        return 0.0
    if a.device != b.device:
        a, b = a.cpu(), b.cpu()
    same = (a == b) | (torch.isnan(a) & torch.isnan(b))     # a work array no sweep has written yet is equal garbage
    if bool(same.all()):
        return 0.0
    d = torch.where(same, torch.zeros_like(a), (a - b).abs())
    if not bool(torch.isfinite(d).all()):
        return float("inf")
    if scale is None:
        fin = torch.isfinite(b)
        scale = float(b[fin].abs().max()) if bool(fin.any()) else 0.0
    return float(d.max()) / (scale + 1e-30)


def absmax(x):
    x = _t(x)
    fin = torch.isfinite(x)
    return float(x[fin].abs().max()) if bool(fin.any()) else 0.0


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


# every per-particle work array of the library, by fetch id; float4 arrays are compared as xyz and w separately
# (w carries a scalar of another scale: warm_start_k / p_past, a gathered payload, rho)
_F1 = {"rho": _lib.F_RHO, "alpha": _lib.F_ALPHA, "rho_derivative": _lib.F_RHO_DERIVATIVE, "rho_adv": _lib.F_RHO_ADV,
       "pressure": _lib.F_PRESSURE, "scalar_a": _lib.F_SCALAR_A, "scalar_b": _lib.F_SCALAR_B, "scalar_c": _lib.F_SCALAR_C}
_F4 = {"vel": _lib.F_FLUID_VEL, "vel_adv": _lib.F_VEL_ADV, "force_a": _lib.F_FORCE_A, "force_b": _lib.F_FORCE_B,
       "vec_a": _lib.F_VEC_A, "vec_b": _lib.F_VEC_B, "payload_2": _lib.F_VEC_C, "payload_1": _lib.F_PAYLOAD_1,
       "payload_3": _lib.F_PAYLOAD_3, "pos_rho": _lib.F_POS_RHO}
# what the generic names hold in each solver (the reference's field names)
ALIASES = {
    "dfsph": {"force_a": "force_ext", "vel.w": "warm_start_k", "payload_2.w": "kappa_j/rho_j (divergence)",
              "payload_3.w": "kappa_j/rho_j (density)", "payload_1.w": "warm_start_k_j/rho_j"},
    "wcsph": {"force_a": "pressure_gradient", "force_b": "viscosity", "vec_a": "tension", "vec_b": "boundary_acc",
              "payload_1.w": "p_j/rho_j^2"},
    "pcisph": {"payload_2": "pos_predict", "vel_adv": "vel_predict", "force_a": "ext_force", "force_b": "press_force",
               "scalar_a": "rho_predict", "scalar_b": "rho_err", "pressure": "press_iter", "payload_1.w": "press_iter_j"},
    "iisph": {"vel_adv": "v_adv", "force_a": "f_adv", "vec_a": "d_ii", "scalar_a": "a_ii", "force_b": "sum_d_ij_p_j",
              "pressure": "p_iter", "scalar_b": "r_sum", "vec_b": "f_press", "vel.w": "p_past", "payload_1.w": "p_j"},
}


def snapshot(ps):
    """every work array of the handle in original particle order, as device tensors"""
    out = {}
    for name, fid in _F1.items():
        out[name] = ps._fetch(fid, 1, torch.float32).reshape(-1)
    for name, fid in _F4.items():
        a = ps._fetch(fid, 4, torch.float32)
        out[name] = a[:, :3]
        out[name + ".w"] = a[:, 3]
    return out


# Residual fields: sums that the solver loop drives towards zero while their TERMS keep their size (D rho / D t =
# sum_j m (v_i - v_j) . grad W_ij: max |.| falls from ~50 to < 1 within the 15 passes).  fp32 summation error is
# eps * sum |terms| whatever the result, so for these fields the error is measured on the un-cancelled scale: the
# largest max |field| the strict run has seen in this step (the first evaluation).  The raw ratio against the
# current max |field| is reported next to it ("... (vs current max)").
RESIDUAL_FIELDS = {"dfsph": ("rho_derivative",)}
# Derived payloads: the stiffness scalars a neighbour contributes, pre-divided by the sweep that produces them --
# kappa_j / rho_j = ((D rho/D t)_j alpha_j / dt) / rho_j (DF:363-367) and ((rho*_j - rho_0) alpha_j / dt^2) / rho_j
# (DF:199-203).  They are pointwise functions of fields that ARE compared between the modes, so comparing them
# between the modes again would only re-measure those inputs, amplified by a particle's alpha or by the cancellation
# in rho* - rho_0.  Each handle's payload is instead checked against its definition, evaluated in float64 from the
# same handle's fields: that pins the production arithmetic (the pre-division) of both modes.
DERIVED_PAYLOADS = {"dfsph": ("payload_2.w", "payload_3.w")}


def _derived_check(err, piece, snaps, stats):
    """dfsph: payload_2.w after a D rho / D t sweep, payload_3.w after a rho* sweep, against their definitions"""
    rec = err.setdefault(piece, {})
    for tag, snap, st in zip(_MODES, snaps, stats):
        rho, alpha = _t(snap["rho"]), _t(snap["alpha"])
        dt = float(np.float32(st.delta_time))
        if "derivative" in piece:
            want = (_t(snap["rho_derivative"]) * alpha / dt) / rho
            got, name = _t(snap["payload_2.w"]), "kappa_j/rho_j == (drho alpha / dt) / rho, " + tag
        else:
            dt2 = float(np.float32(np.float32(dt) * np.float32(dt)))
            want = ((_t(snap["rho_adv"]) - 1000.0) * alpha / dt2) / rho
            got, name = _t(snap["payload_3.w"]), "kappa_j/rho_j == ((rho* - rho_0) alpha / dt^2) / rho, " + tag
        rec[name] = max(rec.get(name, 0.0), relinf(got, want))


def _record(err, piece, solver, snap_f, snap_s, extra=(), scales=None):
    rec = err.setdefault(piece, {})
    al = ALIASES.get(solver, {})
    for k in snap_s:
        if k in DERIVED_PAYLOADS.get(solver, ()):
            continue
        name = al.get(k, k)
        if scales is not None and k in scales:
            raw = relinf(snap_f[k], snap_s[k])
            if raw > 0.0:
                info_name = name + " (vs current max)"
                rec.setdefault("~info", {})[info_name] = max(rec.get("~info", {}).get(info_name, 0.0), raw)
            v = relinf(snap_f[k], snap_s[k], scale=scales[k])
        else:
            v = relinf(snap_f[k], snap_s[k])
        if v > 0.0 or name in rec:
            rec[name] = max(rec.get(name, 0.0), v)
    for name, a, b in extra:
        rec[name] = max(rec.get(name, 0.0), relinf(a, b))


def _mean_where(x, mask):
    x = _t(x)
    return float(x[mask].mean()) if bool(mask.any()) else None


def _avg_check(err, piece, tag, device_value, field, mask, empty_value):
    """device-side average against the float64 statistic of the same handle's field, on the field's scale"""
    host = _mean_where(field, mask)
    host = empty_value if host is None else host
    scale = absmax(field) + 1e-30
    rec = err.setdefault(piece, {})
    rec[tag] = max(rec.get(tag, 0.0), abs(float(device_value) - host) / scale)


_MODES = ("strict", "fast")


def _stat_df_div(err, piece, snaps, stats, first=False):
    for tag, snap, st in zip(_MODES, snaps, stats):
        d = snap["rho_derivative"]
        _avg_check(err, piece, "avg == mean(field), " + tag, st.div_first_err if first else st.div_err, d, d > 0, 0.0)    # DF:274-279


def _stat_df_den(err, piece, snaps, stats):
    for tag, snap, st in zip(_MODES, snaps, stats):
        d = snap["rho_adv"]
        _avg_check(err, piece, "avg == mean(field), " + tag, st.den_err + 1000.0, d, d != 1000.0, 1000.0)                 # DF:141-149


def _stat_pc(err, piece, snaps, stats):
    for tag, snap, st in zip(_MODES, snaps, stats):
        d = snap["scalar_b"]
        _avg_check(err, piece, "avg == mean(field), " + tag, st.pc_err, d, d > 0, 0.0)                                   # PC:123-133


def _stat_ii(err, piece, snaps, stats):
    # II:102-113: mean over p > 0 of a_ii p + r_sum + rho_adv - 1000; a cancelling sum of O(rho_0) terms, so it is
    # compared on the scale of rho_0 -- between the modes and against the float64 statistic of each handle's fields
    rec = err.setdefault(piece, {})
    for tag, snap, st in zip(_MODES, snaps, stats):
        p = _t(snap["pressure"])
        r = _t(snap["scalar_a"]) * p + _t(snap["scalar_b"]) + _t(snap["rho_adv"]) - 1000.0
        host = _mean_where(r, p > 0) or 0.0
        k = "residual == mean(fields), " + tag
        rec[k] = max(rec.get(k, 0.0), abs(float(st.ii_residual) - host) / 1000.0)
    k = "residual, fast vs strict"
    rec[k] = max(rec.get(k, 0.0), abs(float(stats[1].ii_residual) - float(stats[0].ii_residual)) / 1000.0)


def sweeps(solver, ps_s, sol_s, ps_f, sol_f, rigid=False, max_passes=200):
    """One step of `solver` on a strict and a fast handle that hold the SAME caller state, executed one sweep at
    a time; before every sweep the fast handle receives the strict handle's work state.  Returns (err, info):
    err[piece][field] = relinf of the fast output against the strict output (max over the passes of a loop);
    info = iteration counts of both modes, whether every device-side loop decision was identical, error flags."""
    P = _lib
    err, info = {}, {"solver": solver}
    for ps, sol in ((ps_s, sol_s), (ps_f, sol_f)):
        sol.simulate_cnt[None] += 1
        ps.update_grid()
    state = {"flags_equal": True, "scales": {k: 0.0 for k in RESIDUAL_FIELDS.get(solver, ())}}

    def run(piece, phase, stat=None, **kw):
        copy_work_state(ps_f, ps_s)
        ps_s.phase(phase)
        ps_f.phase(phase)
        snaps = (snapshot(ps_s), snapshot(ps_f))
        extra = []
        if rigid:
            extra.append(("rigid_force", ps_f.rigid_particles.force.to_numpy(), ps_s.rigid_particles.force.to_numpy()))
        for k in RESIDUAL_FIELDS.get(solver, ()):
            state["scales"][k] = max(state["scales"][k], absmax(snaps[0][k]))
        _record(err, piece, solver, snaps[1], snaps[0], extra, state["scales"])
        stats = (ps_s.read_stats(), ps_f.read_stats())
        if stat is not None:
            stat(err, piece, snaps, stats, **kw)
            if solver == "dfsph":
                _derived_check(err, piece, snaps, stats)
        for f in ("div_active", "den_active", "loop_active", "div_iters", "den_iters", "pc_iters", "ii_iters"):
            if getattr(stats[0], f) != getattr(stats[1], f):
                state["flags_equal"] = False
        if stats[0].delta_time != stats[1].delta_time:
            err[piece]["delta_time"] = max(err[piece].get("delta_time", 0.0), relinf([stats[1].delta_time], [stats[0].delta_time]))
        return stats[0]

    run("build_lists (rho%s)" % (", alpha" if solver == "dfsph" else ""), P.PH_BUILD_LISTS)
    info["neighbour_counts_equal"] = bool(torch.equal(ps_s.neighbour_counts(), ps_f.neighbour_counts()))
    if solver == "dfsph":
        run("divergence_warm_start", P.PH_DF_WARM_START)
        st = run("derivative_iter_all_rho (first)", P.PH_DF_DRHO_FIRST, _stat_df_div, first=True)
        n = 0
        while st.div_active and n < 15:
            run("divergence_iter_all_vel_adv", P.PH_DF_DIV_VEL)
            st = run("derivative_iter_all_rho", P.PH_DF_DIV_DRHO, _stat_df_div)
            n += 1
        run("ext_force + vel_adv", P.PH_DF_EXT_FORCE_VEL_ADV)
        n = 0
        while True:
            run("compute_all_rho_adv", P.PH_DF_DEN_RHO, _stat_df_den)
            st = run("iter_all_vel_adv", P.PH_DF_DEN_VEL)
            n += 1
            if not st.den_active or n >= max_passes:
                break
        last = P.PH_DF_POSITION
    elif solver == "wcsph":
        run("pressure (Tait)", P.PH_WC_EOS)
        run("pressure gradient + viscosity + tension", P.PH_WC_FORCE)
        last = P.PH_WC_KINEMATIC
    elif solver == "pcisph":
        run("compute_ext_force", P.PH_PC_EXT_FORCE)     # the lists are fresh: only the force sweep runs
        run("predict_vel_pos", P.PH_PC_PREDICT)
        st = run("predict_rho (first)", P.PH_PC_RHO_FIRST, _stat_pc)
        n = 0
        while st.loop_active and n < 80:
            run("iter_press + update_press_force", P.PH_PC_PRESS_FORCE)
            st = run("predict_rho", P.PH_PC_RHO, _stat_pc)
            n += 1
        last = P.PH_PC_INTEGRATION
    elif solver == "iisph":
        run("advection force, v_adv, d_ii", P.PH_II_ADVECT)
        run("rho_adv, a_ii", P.PH_II_AII)
        st = run("pressure_solve begin", P.PH_II_SOLVE_BEGIN)
        n = 0
        while st.loop_active and n < 180:
            run("compute_all_d_ij", P.PH_II_DIJ)
            st = run("update_p", P.PH_II_UPDATE, _stat_ii)
            n += 1
        last = P.PH_II_INTEGRATION
    else:
        raise ValueError("no single-sweep program for solver '%s'" % solver)

    copy_work_state(ps_f, ps_s)
    ps_s.phase(last)
    ps_f.phase(last)
    vs, vf = ps_s._vel4[:ps_s.particle_num], ps_f._vel4[:ps_f.particle_num]
    extra = [("pos", ps_f.fluid_particles.pos.tensor, ps_s.fluid_particles.pos.tensor),
             ("vel", vf[:, :3], vs[:, :3]), (ALIASES[solver].get("vel.w", "vel.w"), vf[:, 3], vs[:, 3])]
    if rigid:
        extra.append(("rigid_force", ps_f.rigid_particles.force.tensor, ps_s.rigid_particles.force.tensor))
    _record(err, "integration", solver, {}, {}, extra)
    ss, sf = ps_s.read_stats(), ps_f.read_stats()
    info["iters"] = {"strict": iters_of(solver, ss), "fast": iters_of(solver, sf)}
    info["loop_flags_equal"] = bool(state["flags_equal"])
    info["error_flags"] = (ss.error_flags, sf.error_flags)
    return err, info


def iters_of(solver, st):
    return {"dfsph": (st.div_iters, st.den_iters), "pcisph": st.pc_iters, "iisph": st.ii_iters, "wcsph": 0}[solver]


def worst(err):
    """(value, 'piece / field') of the largest entry of an error table."""
    best = (0.0, "")
    for piece, rec in err.items():
        for name, v in rec.items():
            if name == "~info":
                continue
            if v >= best[0]:
                best = (v, piece + " / " + name)
    return best


def perturb_velocities_one_ulp(ps, seed=0):
    """Move every velocity component of the caller state by one ulp up or down (seeded)."""
    n = ps.particle_num
    v = ps._vel4[:n, :3].cpu().numpy()
    rng = np.random.default_rng(seed)
    up = rng.integers(0, 2, size=v.shape).astype(bool)
    w = np.where(up, np.nextafter(v, np.float32(np.inf)), np.nextafter(v, np.float32(-np.inf))).astype(np.float32)
    ps._vel4[:n, :3] = torch.from_numpy(w).to(ps._device)


def slab_vs_single(solver="dfsph", steps=3, strict=True, scene_name="small_block", seed=7, cuts=None):
    """Multi-GPU correctness, callable from any torchrun job (tests/mg_worker.py, bench.py --gpus N): every rank
    runs its x-slab of a small jittered scene with large x-velocities (particles cross the cuts and migrate),
    rank 0 also runs the single-domain handle and compares by global particle id.  Returns the verdict on rank 0
    (None elsewhere).  torch.distributed must be initialised with one rank per GPU."""
    import contextlib
    import importlib
    import io

    import torch.distributed as dist

    from . import scene as scene_mod, scenes
    from .ParticleSystem import ParticleSystem

    rank, world = dist.get_rank(), dist.get_world_size()
    cls = getattr(importlib.import_module("cfd_taichi_b200.%s_solver" % solver), "%s_solver" % solver)
    cfg = scenes.shipped(scene_name, solver)
    cfg.pop("solid", None)
    n_global = scene_mod.derive_sizes(cfg)[0]
    rng = np.random.default_rng(seed)
    jit = rng.uniform(-0.008, 0.008, size=(n_global, 3)).astype(np.float32)
    vel = (rng.normal(0, 1.0, size=(n_global, 3)) * np.array([3.0, 0.5, 0.5])).astype(np.float32)
    slab = (rank, world) if cuts is None else (rank, world, cuts)
    with contextlib.redirect_stdout(io.StringIO()):
        ps = ParticleSystem(cfg, strict=strict, solver_name=solver, slab=slab)
        sol = cls(ps, cfg)
    gid0, _, _ = ps.owned_state()
    n0 = len(gid0)
    ps._pos4[:n0, :3] += torch.from_numpy(jit[gid0]).to(ps._device)
    ps._vel4[:n0, :3] = torch.from_numpy(vel[gid0]).to(ps._device)
    hist = []
    for _ in range(steps):
        sol.step()
        info = ps.comm_info()
        hist.append((info["owned"], info["ghosts"]))
    st = sol.stats()
    gid, pos, v4 = ps.owned_state()
    out = dict(rank=rank, gid0=gid0, gid=gid, pos=pos, vel=v4, hist=hist, flags=st.error_flags,
               iters=(st.div_iters, st.den_iters, st.pc_iters, st.ii_iters))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(out, gathered, dst=0)
    res = None
    if rank == 0:
        with contextlib.redirect_stdout(io.StringIO()):
            ps1 = ParticleSystem(cfg, strict=strict, solver_name=solver)
            sol1 = cls(ps1, cfg)
        ps1._pos4[:n_global, :3] += torch.from_numpy(jit).to(ps1._device)
        ps1._vel4[:n_global, :3] = torch.from_numpy(vel).to(ps1._device)
        for _ in range(steps):
            sol1.step()
        st1 = sol1.stats()
        ref_pos, ref_vel = ps1._pos4[:n_global, :3].cpu().numpy(), ps1._vel4[:n_global].cpu().numpy()
        ps1.close()
        gids = np.concatenate([g["gid"] for g in gathered])
        pos = np.concatenate([g["pos"] for g in gathered])
        vel4 = np.concatenate([g["vel"] for g in gathered])
        perm_ok = bool(np.array_equal(np.sort(gids), np.arange(n_global)))
        owner0 = np.full(n_global, -1)
        owner1 = np.full(n_global, -1)
        for g in gathered:
            owner0[g["gid0"]] = g["rank"]
            owner1[g["gid"]] = g["rank"]
        exact = False
        dpos = dvel = float("nan")
        if perm_ok:
            order = np.argsort(gids)
            pos, vel4 = pos[order], vel4[order]
            exact = bool(np.array_equal(pos, ref_pos) and np.array_equal(vel4, ref_vel))
            dpos, dvel = float(np.abs(pos - ref_pos).max()), float(np.abs(vel4 - ref_vel).max())
        ref_iters = (st1.div_iters, st1.den_iters, st1.pc_iters, st1.ii_iters)
        res = dict(solver=solver, scene=scene_name, particles=int(n_global), ranks=world, steps=steps,
                   kernels="strict-fp32" if strict else "fast-fp32",
                   slab_vs_single_bit_exact=exact, every_particle_owned_once=perm_ok,
                   iters_ok=bool(all(tuple(g["iters"]) == ref_iters for g in gathered)),
                   migrated_particles=int((owner0 != owner1).sum()),
                   crossed_two_cuts=int((np.abs(owner0 - owner1) >= 2).sum()),
                   max_abs_dpos=dpos, max_abs_dvel=dvel, error_flags=[int(g["flags"]) for g in gathered],
                   owned_ghosts_last_step=[list(g["hist"][-1]) for g in gathered])
        res["ok"] = bool(perm_ok and res["iters_ok"] and exact and not any(res["error_flags"]))   # both modes: bit for bit
    ps.close()
    dist.barrier()
    return res
