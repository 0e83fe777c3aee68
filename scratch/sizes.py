"""Single-GPU timing of the other BASELINE configs: DFSPH 8 M, PCISPH / IISPH 4 M, WCSPH 1 M (developer aid)."""
import sys, time, torch, contextlib, io
sys.path.insert(0, '.')
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200 import scenes
import importlib
def run(solver, n_side, steps, warm=3):
    cfg = scenes.breaking_dam(n_side, solver=solver) if 'solver' in scenes.breaking_dam.__code__.co_varnames else scenes.breaking_dam(n_side)
    cfg['solver']['name'] = solver
    if solver == 'pcisph': cfg['solver']['delta_time'] = 1.5e-4
    if solver in ('iisph', 'wcsph'): cfg['solver']['delta_time'] = 2.5e-4
    with contextlib.redirect_stdout(io.StringIO()):
        ps = ParticleSystem(cfg, strict=False, solver_name=solver)
        mod = importlib.import_module('cfd_taichi_b200.%s_solver' % solver)
        sol = getattr(mod, '%s_solver' % solver)(ps, cfg)
    for _ in range(warm): sol.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): sol.step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = sol.stats()
    print(solver, 'N', ps.particle_num, 'ms/step %.3f' % ms, 'Mps/s %.1f' % (ps.particle_num / ms / 1e3),
          'iters', st.div_iters, st.den_iters, getattr(st, 'pc_iters', None), getattr(st, 'ii_iters', None), 'flags', st.error_flags,
          'mem GB %.1f' % (torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9), flush=True)
    ps.close()
for a in sys.argv[1:]:
    f = a.split(':')       # solver:n_side:steps[:warm-up steps]
    run(f[0], int(f[1]), int(f[2]), int(f[3]) if len(f) > 3 else 3)
