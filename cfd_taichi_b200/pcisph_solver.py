"""pcisph_solver -- drop-in mirror of the reference class (pcisph_solver.py:5-240): delta precompute,
predict / correct pressure iteration with the residual and the loop decision kept on the device."""
import ctypes

from . import _lib
from .fields import FetchedField, HostScalar
from .solver_base import solver_base


class pcisph_solver(solver_base):
    solver_name = 'pcisph'

    def __init__(self, particle_system, config):
        super(pcisph_solver, self).__init__(particle_system, config)
        self.pos_predict = FetchedField(self.ps, _lib.F_VEC_C, 3)                    # PC:9-17
        self.vel_predict = FetchedField(self.ps, _lib.F_VEL_ADV, 3)
        self.ext_force = FetchedField(self.ps, _lib.F_FORCE_A, 3)
        self.press_force = FetchedField(self.ps, _lib.F_FORCE_B, 3)
        self.rho_predict = FetchedField(self.ps, _lib.F_SCALAR_A)
        self.rho_err = FetchedField(self.ps, _lib.F_SCALAR_B)
        self.press_iter = FetchedField(self.ps, _lib.F_PRESSURE)
        self.rho_max_err_percent = .1                                                # PC:19-21
        self.min_iteration = 1
        self.max_iteration = 80
        dt = float(config.get('solver').get('delta_time'))
        import numpy as np
        dtf = float(np.float32(dt))
        self.beta = dtf * dtf * self.ps.particle_m * self.ps.particle_m * 2 / (self.rho_0 ** 2)   # PC:23
        self.delta = HostScalar(0.0)
        self.pre_compute()                                                           # PC:26

    def pre_compute(self):                                                           # PC:28-37
        ps = self.ps
        _lib.check(self._lib.sph_pcisph_precompute(ps._h, ps._stream()), ps._h)
        max_index = ps.get_max_neighbor_particle_index()
        if max_index >= 0:
            _lib.check(self._lib.sph_pcisph_delta(ps._h, int(max_index), ps._stream()), ps._h)
        st = self.stats()
        self.delta[None] = st.pc_delta
        print('PCISPH parameter delta: {}, beta: {}'.format(self.delta[None], self.beta))

    def compute_ext_force(self):                                                     # PC:220-226
        self.ps.phase(_lib.PH_PC_EXT_FORCE)

    def iteration(self):                                                             # PC:47-70
        self.ps.phase(_lib.PH_PC_ITERATION)
        if self.verbose:
            s = self.stats()
            print('\t\tIter cnt: {}, error: {}'.format(s.pc_iters, s.pc_err))

    def integration(self):                                                           # PC:200-218
        self.ps.phase(_lib.PH_PC_INTEGRATION)

    def step(self):                                                                  # PC:233-240
        self._full_step(1)
        if self.verbose:
            s = self.stats()
            print('\t\tIter cnt: {}, error: {}'.format(s.pc_iters, s.pc_err))
