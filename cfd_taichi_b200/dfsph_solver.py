"""dfsph_solver -- drop-in mirror of the reference class (dfsph_solver.py:5-445).

step() = solver_base.step + initialize + iterate, executed as one stream-ordered sequence of CUDA
kernels; the divergence-free and constant-density loops keep their average-error reductions and
their continue/stop decisions on the device (no host sync per iteration).
"""
from . import _lib
from .fields import FetchedField
from .solver_base import solver_base


class dfsph_solver(solver_base):
    solver_name = 'dfsph'

    def __init__(self, particle_system, config):
        super(dfsph_solver, self).__init__(particle_system, config)
        self.alpha = FetchedField(self.ps, _lib.F_ALPHA)                             # DF:10-17
        self.rho_adv = FetchedField(self.ps, _lib.F_RHO_ADV)
        self.rho_derivative = FetchedField(self.ps, _lib.F_RHO_DERIVATIVE)
        self.vel_adv = FetchedField(self.ps, _lib.F_VEL_ADV, 3)
        self.force_ext = FetchedField(self.ps, _lib.F_FORCE_A, 3)
        self.warm_start_k = _WarmStartK(self.ps)
        self.min_iteration_density = 2                                               # DF:21-29
        self.density_threshold = 0.1
        self.min_iteration_density_divergence = 1
        self.max_iteration_density_divergence = 15
        self.density_divergence_threshold = 10
        self.warm_start = True
        self.adaptive_dt = True
        self.max_dt = 1e-3
        self.min_dt = 1e-5

    # piecewise API (DF:423-438) ------------------------------------------------------------------
    def initialize(self):                                                            # DF:423-426
        self.ps.phase(_lib.PH_DF_INITIALIZE)

    def correct_divergence_error(self):                                              # DF:393-416
        self.ps.phase(_lib.PH_DF_DIVERGENCE)
        if self.verbose:
            s = self.stats()
            print('[divergence iteration] count: {}, first error {}, error {}'.format(
                s.div_iters, s.div_first_err, s.div_err))

    def compute_all_ext_force(self):                                                 # DF:91-96
        self._ext_pending = True

    def compute_all_vel_adv(self):                                                   # DF:98-122
        # tension + viscosity + f_ext + v* + adaptive dt are one fused pass
        self.ps.phase(_lib.PH_DF_EXT_FORCE_VEL_ADV)
        self._ext_pending = False

    def correct_density_error(self):                                                 # DF:221-233
        self.ps.phase(_lib.PH_DF_DENSITY)
        if self.verbose:
            s = self.stats()
            print('[density iteration] count: {}, error {}'.format(s.den_iters, s.den_err))

    def compute_all_position(self):                                                  # DF:235-250
        self.ps.phase(_lib.PH_DF_POSITION)

    def iterate(self):                                                               # DF:428-438
        self.correct_divergence_error()
        self.compute_all_ext_force()
        self.compute_all_vel_adv()
        self.correct_density_error()
        self.compute_all_position()

    def step(self):                                                                  # DF:440-445
        self._full_step(1)
        if self.verbose:
            s = self.stats()
            print('[divergence iteration] count: {}, first error {}, error {}'.format(
                s.div_iters, s.div_first_err, s.div_err))
            print('[density iteration] count: {}, error {}'.format(s.den_iters, s.den_err))


class _WarmStartK:
    """warm_start_k (DF:17) rides in the .w lane of the velocity float4."""

    def __init__(self, ps):
        self._ps = ps

    def to_numpy(self):
        return self._ps._vel4[:self._ps.particle_num, 3].contiguous().cpu().numpy()

    def from_numpy(self, arr):
        import torch
        self._ps._vel4[:self._ps.particle_num, 3] = torch.as_tensor(arr, dtype=torch.float32,
                                                                     device=self._ps._device)
