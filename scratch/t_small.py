"""ms per step of the small shipped scenes (launch-bound): usage t_small.py solver scene steps"""
import contextlib, importlib, io, sys, torch
sys.path.insert(0, '.')
from cfd_taichi_b200 import scenes
from cfd_taichi_b200.ParticleSystem import ParticleSystem
for arg in sys.argv[1:]:
    solver, scene, steps = arg.split(':'); steps = int(steps)
    cfg = scenes.shipped(scene, solver); cfg.pop('solid', None)
    with contextlib.redirect_stdout(io.StringIO()):
        ps = ParticleSystem(cfg, solver_name=solver)
        sol = getattr(importlib.import_module('cfd_taichi_b200.%s_solver' % solver), '%s_solver' % solver)(ps, cfg)
    for _ in range(20): sol.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): sol.step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(solver, scene, 'N', ps.particle_num, 'ms/step %.4f' % ms, 'Mps/s %.1f' % (ps.particle_num / ms / 1e3), 'flags', sol.stats().error_flags, flush=True)
    ps.close()
