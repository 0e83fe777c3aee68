"""wcsph_solver -- drop-in mirror of the reference class (wcsph_solver.py:5-144): Tait equation of
state, symmetric pressure gradient, Akinci boundary pressure, viscosity, tension, symplectic Euler."""
from . import _lib
from .fields import FetchedField
from .solver_base import solver_base


class wcsph_solver(solver_base):
    solver_name = 'wcsph'

    def __init__(self, particle_system, config):
        super(wcsph_solver, self).__init__(particle_system, config)
        self.pressure = FetchedField(self.ps, _lib.F_PRESSURE)                       # WC:10-15
        self.pressure_gradient = FetchedField(self.ps, _lib.F_FORCE_A, 3)
        self.viscosity = FetchedField(self.ps, _lib.F_FORCE_B, 3)
        self.tension = FetchedField(self.ps, _lib.F_VEC_A, 3)
        self.boundary_acc = FetchedField(self.ps, _lib.F_VEC_B, 3)
        self.viscosity_epsilon = 0.01                                                # WC:17-22
        self.viscosity_c_s = 10
        self.viscosity_alpha = 0.08
        self.tension_k = 0.2
        self.gamma = 7
        self.B = 70000

    def pressure_phase(self):                                                        # WC:32-38
        self.ps.phase(_lib.PH_WC_PRESSURE)

    def kinematic_phase(self):                                                       # WC:40-63
        self.ps.phase(_lib.PH_WC_KINEMATIC)

    def step(self):                                                                  # WC:25-30
        self._full_step(1)
