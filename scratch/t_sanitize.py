"""Small scenes for compute-sanitizer (scratch/sanitize.sh): a few steps of every solver on the 5.9 k-particle block,
fast and strict kernels, plus the coupled rigid scene; slab mode when launched under torchrun."""
import contextlib, io, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from cfd_taichi_b200 import scene, scenes, selfcheck
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.rigid_solver import rigid_solver
import importlib

def solver_cls(name):
    return getattr(importlib.import_module('cfd_taichi_b200.%s_solver' % name), '%s_solver' % name)

what = sys.argv[1] if len(sys.argv) > 1 else 'solvers'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if what == 'slab':
    import torch.distributed as dist
    rank = int(os.environ['RANK']); torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', rank))))
    res = selfcheck.slab_vs_single(solver='dfsph', steps=steps, strict=True)
    if rank == 0: print('slab', res['ok'], res['migrated_particles'], res['error_flags'])
    dist.destroy_process_group()
    sys.exit(0)
if what == 'solvers':
    for name in ('dfsph', 'wcsph', 'pcisph', 'iisph', 'pbf'):
        for strict in (True, False):
            cfg = scenes.shipped('small_block', name)
            with contextlib.redirect_stdout(io.StringIO()):
                ps = ParticleSystem(cfg, strict=strict, solver_name=name); sol = solver_cls(name)(ps, cfg)
            n = ps.particle_num
            rng = np.random.default_rng(1)
            ps._vel4[:n, :3] = torch.from_numpy(rng.normal(0, 0.5, size=(n, 3)).astype(np.float32)).to(ps._device)
            for _ in range(steps): sol.step()
            st = sol.stats(); print(name, 'strict' if strict else 'fast', 'flags', st.error_flags, 'launches', st.kernel_launches, flush=True)
            ps.close()
if what == 'rigid':
    for strict in (True, False):
        cfg = scenes.make_scene([2.0, 2.0, 1.0], [0.1, 0.1, 0.1], [0.6, 0.8, 0.8], 'dfsph', 1e-4,
                                solid={'mesh': './obj/cube1.STL', 'voxel_radius': 0.025, 'rho_0': 2000, 'scale': 0.5,
                                       'pos_offset': [0.85, 0.0, 0.2], 'attitude_offset': [0.0, 0.0, 0.0], 'fill': True, 'active': True})
        with contextlib.redirect_stdout(io.StringIO()):
            ps = ParticleSystem(cfg, strict=strict, solver_name='dfsph'); sol = solver_cls('dfsph')(ps, cfg); rs = rigid_solver(ps, cfg)
        for _ in range(steps): sol.step(); rs.step()
        st = sol.stats(); print('rigid', 'strict' if strict else 'fast', 'flags', st.error_flags, 'rigid particles', ps.rigid_particles_num, flush=True)
        ps.close()
