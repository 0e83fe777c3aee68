import contextlib, io, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from cfd_taichi_b200 import scenes, _lib
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.dfsph_solver import dfsph_solver
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
cfg = scenes.shipped("small_block", "dfsph"); n_global = 5879
rng = np.random.default_rng(7)
jit = rng.uniform(-0.008, 0.008, size=(n_global, 3)).astype(np.float32)
vel = (rng.normal(0, 1.0, size=(n_global, 3)) * np.array([3.0, 0.5, 0.5])).astype(np.float32)
with contextlib.redirect_stdout(io.StringIO()):
    ps = ParticleSystem(cfg, strict=True, solver_name="dfsph", slab=(rank, world)); sol = dfsph_solver(ps, cfg)
    if rank == 0:
        ps1 = ParticleSystem(cfg, strict=True, solver_name="dfsph"); sol1 = dfsph_solver(ps1, cfg)
gid, pos, v4 = ps.owned_state(); n0 = len(gid)
ps._pos4[:n0, :3] += torch.from_numpy(jit[gid]).to(ps._device); ps._vel4[:n0, :3] = torch.from_numpy(vel[gid]).to(ps._device)
if rank == 0:
    ps1._pos4[:n_global, :3] += torch.from_numpy(jit).to(ps1._device); ps1._vel4[:n_global, :3] = torch.from_numpy(vel).to(ps1._device)
def gather(name, arr_local, gids):
    out = [None]*world if rank == 0 else None
    dist.gather_object((gids, arr_local), out, dst=0)
    if rank == 0:
        g = np.concatenate([a for a, _ in out]); v = np.concatenate([b for _, b in out]); o = np.argsort(g)
        return g[o], v[o]
    return None, None
def owned(field_t, width=1):
    info = ps.comm_info(); n = info['owned']
    return field_t[:n].cpu().numpy()
def cmp(tag, local, ref):
    info = ps.comm_info(); n = info['owned']
    g, v = gather(tag, local, ps._gid[:n].cpu().numpy())
    if rank == 0:
        d = np.abs(v - ref[g] if ref.ndim == v.ndim else v - ref[g])
        print(tag, 'maxdiff %.3e' % d.max(), 'n_bad', int((d.reshape(len(g), -1).max(1) > 0).sum()), flush=True)
        bad = np.nonzero(d.reshape(len(g), -1).max(1) > 0)[0]
        if len(bad): print('   bad gids', g[bad][:10], 'x', ps1._pos4[g[bad][:10], 0].cpu().numpy())
for step in range(3):
    ps.update_grid()
    if rank == 0: ps1.update_grid()
    info = ps.comm_info(); print('rank', rank, 'step', step, info, flush=True)
    sol.initialize()
    if rank == 0: sol1.initialize()
    ref = lambda f: f.to_numpy() if rank == 0 else None
    cmp('rho', owned(sol.rho.to_torch()), sol1.rho.to_numpy() if rank == 0 else None)
    cmp('alpha', owned(sol.alpha.to_torch()), sol1.alpha.to_numpy() if rank == 0 else None)
    cmp('nbr', owned(ps.neighbour_counts()), ps1.neighbour_counts().cpu().numpy() if rank == 0 else None)
    sol.correct_divergence_error()
    if rank == 0: sol1.correct_divergence_error()
    st = sol.stats(); print('rank', rank, 'div', st.div_iters, st.div_first_err, st.div_err, (sol1.stats().div_iters, sol1.stats().div_first_err, sol1.stats().div_err) if rank == 0 else '', flush=True)
    cmp('drho', owned(sol.rho_derivative.to_torch()), sol1.rho_derivative.to_numpy() if rank == 0 else None)
    cmp('vel_after_div', owned(ps._fetch(_lib.F_FLUID_VEL, 4, torch.float32)), ps1._fetch(_lib.F_FLUID_VEL, 4, torch.float32).cpu().numpy() if rank == 0 else None)
    sol.compute_all_ext_force(); sol.compute_all_vel_adv()
    if rank == 0: sol1.compute_all_ext_force(); sol1.compute_all_vel_adv()
    cmp('vel_adv', owned(sol.vel_adv.to_torch()), sol1.vel_adv.to_numpy() if rank == 0 else None)
    sol.correct_density_error()
    if rank == 0: sol1.correct_density_error()
    cmp('rho_adv', owned(sol.rho_adv.to_torch()), sol1.rho_adv.to_numpy() if rank == 0 else None)
    cmp('vel_adv2', owned(sol.vel_adv.to_torch()), sol1.vel_adv.to_numpy() if rank == 0 else None)
    sol.compute_all_position()
    if rank == 0: sol1.compute_all_position()
    g_, p_, v_ = ps.owned_state()
    cmp('pos', p_, ps1._pos4[:n_global, :3].cpu().numpy() if rank == 0 else None)
dist.barrier(); dist.destroy_process_group()
