import sys, contextlib, io
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from cfd_taichi_b200 import scenes, selfcheck, _lib
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.iisph_solver import iisph_solver
cfg = scenes.shipped("small_block", "iisph")
def mk(strict):
    with contextlib.redirect_stdout(io.StringIO()):
        ps = ParticleSystem(cfg, strict=strict, solver_name="iisph"); return ps, iisph_solver(ps, cfg)
ps_s, sol_s = mk(True); ps_f, sol_f = mk(False)
for step in range(170):
    if step in (100, 101, 120, 140, 160):
        selfcheck.copy_caller_state(ps_f, sol_f, ps_s, sol_s)
        # instrumented walk: print residuals of both modes after each update
        orig = selfcheck._stat_ii
        def stat(err, piece, snaps, stats):
            print("   residual strict %.9g fast %.9g  iters %d/%d active %d/%d" % (stats[0].ii_residual, stats[1].ii_residual, stats[0].ii_iters, stats[1].ii_iters, stats[0].loop_active, stats[1].loop_active))
            orig(err, piece, snaps, stats)
        selfcheck._stat_ii = stat
        err, info = selfcheck.sweeps("iisph", ps_s, sol_s, ps_f, sol_f)
        selfcheck._stat_ii = orig
        print("step", step, info, selfcheck.worst(err))
        for k, v in err.items(): print("   %-40s %s" % (k, {n: "%.1e" % x for n, x in v.items() if not isinstance(x, dict)}))
    else:
        sol_s.step()
