"""Per-phase times of the e2e loop (upload / step / download), CUDA events on the caller's stream."""
import contextlib, ctypes, io, sys, time, torch
sys.path.insert(0, '.')
from cfd_taichi_b200 import _lib, scenes
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.dfsph_solver import dfsph_solver
cfg = scenes.breaking_dam(100)
with contextlib.redirect_stdout(io.StringIO()):
    ps = ParticleSystem(cfg, solver_name='dfsph'); sol = dfsph_solver(ps, cfg)
for _ in range(5): sol.step()
n = ps.particle_num
hp = torch.empty((n, 3), dtype=torch.float32).pin_memory(); hv = torch.empty((n, 3), dtype=torch.float32).pin_memory()
L, h, s = ps._lib, ps._h, ps._stream()
_lib.check(L.sph_download_state_xyz(h, hp.data_ptr(), hv.data_ptr(), s), h)
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(12)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(12):
    ev[k][0].record()
    _lib.check(L.sph_upload_state_xyz(h, hp.data_ptr(), hv.data_ptr(), s), h)
    ev[k][1].record()
    _lib.check(L.sph_step(h, 1, s), h)
    ev[k][2].record()
    _lib.check(L.sph_download_state_xyz(h, hp.data_ptr(), hv.data_ptr(), s), h)
    ev[k][3].record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 12 * 1e3
up = sum(e[0].elapsed_time(e[1]) for e in ev[2:]) / 10; st = sum(e[1].elapsed_time(e[2]) for e in ev[2:]) / 10; dn = sum(e[2].elapsed_time(e[3]) for e in ev[2:]) / 10
gap = sum(ev[k][3].elapsed_time(ev[k + 1][0]) for k in range(2, 11)) / 9
print('wall ms/step %.3f | upload (stream part) %.3f  step %.3f  download %.3f  gap to next %.3f' % (wall, up, st, dn, gap))
