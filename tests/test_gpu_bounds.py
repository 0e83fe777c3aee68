"""Index checks of our own instead of compute-sanitizer (closed on the GPU pool this repository is developed on).

cfd_taichi_b200/_build_debug/libsph_b200.so is the library compiled with -DSPH_DEBUG_BOUNDS=1: every neighbour-list
entry, list length and candidate segment is validated before use and SPH_ERR_BOUNDS (bit 6 of error_flags) is
latched on a violation.  Every solver (both arithmetic modes), the coupled rigid scene and a piecewise walk run
under it on small scenes -- in a subprocess, because the library path is fixed at import time -- and a deliberately
corrupted list entry shows that the checks are live."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import contextlib, ctypes, importlib, io, sys
import numpy as np, torch
sys.path.insert(0, %(root)r)
from cfd_taichi_b200 import _lib, scenes, selfcheck
from cfd_taichi_b200.ParticleSystem import ParticleSystem
from cfd_taichi_b200.rigid_solver import rigid_solver
assert _lib.LIB_PATH.endswith('_build_debug/libsph_b200.so'), _lib.LIB_PATH
def make(cfg, name, strict):
    with contextlib.redirect_stdout(io.StringIO()):
        ps = ParticleSystem(cfg, strict=strict, solver_name=name, base_dir=%(root)r)
        sol = getattr(importlib.import_module('cfd_taichi_b200.%%s_solver' %% name), '%%s_solver' %% name)(ps, cfg)
    return ps, sol
worst = 0
for name in ('dfsph', 'wcsph', 'pcisph', 'iisph', 'pbf'):
    for strict in (True, False):
        cfg = scenes.shipped('small_block', name)
        ps, sol = make(cfg, name, strict)
        n = ps.particle_num
        rng = np.random.default_rng(1)
        ps._vel4[:n, :3] = torch.from_numpy(rng.normal(0, 0.5, size=(n, 3)).astype(np.float32)).to(ps._device)
        for _ in range(4): sol.step()
        f = sol.stats().error_flags
        print(name, 'strict' if strict else 'fast', 'flags', f, flush=True)
        worst |= f
        ps.close()
for strict in (True, False):
    cfg = scenes.shipped('dam_flush_cube', 'dfsph')
    ps, sol = make(cfg, 'dfsph', strict)
    rs = rigid_solver(ps, cfg)
    for _ in range(3): sol.step(); rs.step()
    f = sol.stats().error_flags
    print('rigid', 'strict' if strict else 'fast', 'flags', f, flush=True)
    worst |= f
    ps.close()
# the sweep-by-sweep walk (every single-sweep phase + sph_copy_work_state)
for name in ('dfsph', 'iisph', 'pcisph', 'wcsph'):
    cfg = scenes.shipped('small_block', name)
    ps_s, sol_s = make(cfg, name, True); ps_f, sol_f = make(cfg, name, False)
    for _ in range(2): sol_s.step()
    selfcheck.copy_caller_state(ps_f, sol_f, ps_s, sol_s)
    err, info = selfcheck.sweeps(name, ps_s, sol_s, ps_f, sol_f)
    print('walk', name, info['error_flags'], flush=True)
    worst |= info['error_flags'][0] | info['error_flags'][1]
    ps_s.close(); ps_f.close()
print('CLEAN' if worst == 0 else 'FLAGS %%d' %% worst)
# the checks are live: corrupt one list entry between the list build and the first sweep
cfg = scenes.shipped('small_block', 'dfsph')
for strict in (True, False):
    ps, sol = make(cfg, 'dfsph', strict)
    sol.simulate_cnt[None] += 1
    ps.update_grid()
    ps.phase(_lib.PH_BUILD_LISTS)
    L = ps._lib
    L.sph_debug_poke_list.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint]
    assert L.sph_debug_poke_list(ps._h, 1000, 0, 0x00ffffff) == 0
    # (the strict DFSPH solve sweeps take j from the gradient records; the force sweep walks the list in both modes)
    ps.phase(_lib.PH_DF_WARM_START); ps.phase(_lib.PH_DF_DRHO_FIRST); ps.phase(_lib.PH_DF_EXT_FORCE_VEL_ADV)
    f = ps.read_stats().error_flags
    print('poked', 'strict' if strict else 'fast', 'flags', f, flush=True)
    assert f & 64, f
    ps.close()
print('LIVE')
"""


def test_bounds_checked_build_is_clean_and_live(built):
    from cfd_taichi_b200 import build as cuda_build
    lib = cuda_build.DEBUG_LIB
    assert os.path.exists(lib), "bounds-checked library missing: __graft_entry__.build() builds it"
    env = dict(os.environ, SPH_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, "stdout:\n%s\nstderr:\n%s" % (r.stdout[-2500:], r.stderr[-2500:])
    assert "CLEAN" in r.stdout and "LIVE" in r.stdout, r.stdout[-3000:]
