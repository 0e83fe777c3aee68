"""iisph_solver -- drop-in mirror of the reference class (iisph_solver.py:5-349): d_ii / a_ii
precompute, relaxed-Jacobi pressure solve (omega = 0.5) with the residual and the divergence-trend
break evaluated on the device, pressure-force integration."""
from . import _lib
from .fields import FetchedField
from .solver_base import solver_base


class iisph_solver(solver_base):
    solver_name = 'iisph'

    def __init__(self, particle_system, config):
        super(iisph_solver, self).__init__(particle_system, config)
        self.v_adv = FetchedField(self.ps, _lib.F_VEL_ADV, 3)                        # II:10-24
        self.f_adv = FetchedField(self.ps, _lib.F_FORCE_A, 3)
        self.d_ii = FetchedField(self.ps, _lib.F_VEC_A, 3)
        self.a_ii = FetchedField(self.ps, _lib.F_SCALAR_A)
        self.d_ij = FetchedField(self.ps, _lib.F_FORCE_B, 3)
        self.rho_adv = FetchedField(self.ps, _lib.F_RHO_ADV)
        self.p_iter = FetchedField(self.ps, _lib.F_PRESSURE)
        self.p_past = _PPast(self.ps)
        self.r_sum = FetchedField(self.ps, _lib.F_SCALAR_B)
        self.f_press = FetchedField(self.ps, _lib.F_VEC_B, 3)
        self.omega = 0.5                                                             # II:26-29
        self.max_iter_cnt = 180
        self.min_iter_cnt = 1
        self.rho_err_percent = .1

    def predict_advection(self):                                                     # II:35-75
        self.ps.phase(_lib.PH_II_PREDICT_ADVECTION)

    def pressure_solve(self):                                                        # II:78-100
        self.ps.phase(_lib.PH_II_PRESSURE_SOLVE)
        if self.verbose:
            s = self.stats()
            print("Iter cnt: ", s.ii_iters, s.ii_residual)

    def intergation(self):                                                           # II:184-206 (sic)
        self.ps.phase(_lib.PH_II_INTEGRATION)

    integration = intergation

    def step(self):                                                                  # II:342-349
        self._full_step(1)
        if self.verbose:
            s = self.stats()
            print("Iter cnt: ", s.ii_iters, s.ii_residual)


class _PPast:
    """p_past (II:21) rides in the .w lane of the velocity float4."""

    def __init__(self, ps):
        self._ps = ps

    def to_numpy(self):
        return self._ps._vel4[:self._ps.particle_num, 3].contiguous().cpu().numpy()
