"""Warm a solver up, then run a few steps between cudaProfilerStart / Stop (ncu --profile-from-start off).
usage: t_prof_solver.py solver n_side|scene warm steps [strict]"""
import contextlib, importlib, io, sys
import torch
sys.path.insert(0, '.')
from cfd_taichi_b200 import scenes
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.rigid_solver import rigid_solver
solver, what, warm, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
strict = len(sys.argv) > 5 and sys.argv[5] == 'strict'
DT = {'dfsph': 1e-3, 'pcisph': 1.5e-4, 'iisph': 2.5e-4, 'wcsph': 2.5e-4, 'pbf': 2.5e-4}
cfg = scenes.breaking_dam(int(what), solver, DT[solver]) if what.isdigit() else scenes.shipped(what, solver)
with contextlib.redirect_stdout(io.StringIO()):
    ps = ParticleSystem(cfg, strict=strict, solver_name=solver)
    sol = getattr(importlib.import_module('cfd_taichi_b200.%s_solver' % solver), '%s_solver' % solver)(ps, cfg)
    rs = rigid_solver(ps, cfg) if cfg.get('solid') else None
def step():
    sol.step()
    if rs: rs.step()
for _ in range(warm): step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps): step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
st = sol.stats()
print(solver, what, 'N', ps.particle_num, 'iters', st.div_iters, st.den_iters, st.pc_iters, st.ii_iters, 'flags', st.error_flags)
