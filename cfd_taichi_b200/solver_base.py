"""solver_base -- drop-in mirror of the reference class (solver_base.py:4-245).

Holds the solver-level scalars and result fields; the sweeps themselves (compute_all_rho SB:41-72,
solve_all_viscosity SB:170-202, solve_all_tension SB:204-217 and the SPH kernels SB:74-129) are
fused into the CUDA passes of each concrete solver (csrc/sph_sweeps.cu).
"""
import ctypes

import torch

from . import _lib
from .fields import DeviceScalar, FetchedField, HostScalar


class solver_base:
    solver_name = None   # set by subclasses; main.py:65-68 resolves "<name>_solver"

    def __init__(self, particle_system, config):
        particle_count = particle_system.particle_num
        scene_config = config.get('scene')
        solver_config = config.get('solver')
        self.particle_count = particle_count
        self.ps = particle_system
        if self.solver_name is not None:
            self.ps._ensure_solver(self.solver_name)
        self._lib = particle_system._lib
        self.rho = FetchedField(self.ps, _lib.F_RHO)                                 # SB:14
        self.delta_time = DeviceScalar(self._get_dt, self._set_dt)                   # SB:15-16
        self.kernel_h = self.ps.particle_radius * 4                                  # SB:17
        self.v_decay_proportion = 0.5
        self.rho_0 = 1000
        self.gravity = scene_config.get('gravity')
        self.simulate_cnt = HostScalar(0)                                            # SB:21
        self.viscosity_epsilon = 0.01                                                # SB:23-26
        self.viscosity_c_s = 13
        self.viscosity_alpha = 0.08
        self.tension_k = 0.5
        self.boundary_handle = 1 if solver_config.get('boundary_handle', True) else 0   # SB:31-35
        self.fs_couple = 1 if solver_config.get('fs_couple', True) else 0
        self.two_way_couple = 1
        self.clamp_boundary_handle = 0
        self.akinci2012_boundary_handle = 1
        self.artificial_friction = 0.9999                                            # SB:37
        self.verbose = bool(solver_config.get('verbose', False))
        print("\033[32m[Solver]: {}\033[0m".format(solver_config.get('name')))

    # -- delta_time[None] lives in the device control block (DFSPH rewrites it every step) ---------
    def _get_dt(self):
        return float(self.ps.read_stats().delta_time)

    def _set_dt(self, value):
        _lib.check(self._lib.sph_set_delta_time(self.ps._h, ctypes.c_float(value), self.ps._stream()), self.ps._h)

    def reset(self):                                                                 # SB:131-134
        pass

    def step(self):                                                                  # SB:136-143
        """Counter + grid rebuild (+ reset).  Concrete solvers call the fused sph_step instead, which
        does the same prologue on the device; this entry exists for piecewise drivers."""
        self.simulate_cnt[None] += 1
        self.ps.reset_grid()
        self.ps.update_grid()
        self.reset()

    def visualize_rho(self):                                                         # SB:219-232
        """ps.rgb[i] = (0, 0.28, (rho_i - min) / (max - min)); reductions and colour map on the device."""
        self._visualize(0)

    def visualize_neighbour(self):                                                   # SB:234-245
        self._visualize(1)

    def _visualize(self, what):
        rgb = self.ps.rgb.tensor
        _lib.check(self._lib.sph_visualize(self.ps._h, what, rgb.data_ptr(), rgb.shape[1], rgb.shape[0],
                                           self.ps._stream()), self.ps._h)

    def _full_step(self, n=1):
        self.simulate_cnt[None] += n
        _lib.check(self._lib.sph_step(self.ps._h, n, self.ps._stream()), self.ps._h)

    def stats(self):
        return self.ps.read_stats()
