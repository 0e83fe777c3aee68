"""pbf_solver -- mirror of the reference class (pbf_solver.py:6-186): position-based fluids, lambda /
delta-p sweeps with the poly6 and spiky kernels, artificial pressure s_corr, XSPH viscosity.

The reference class cannot compile at its HEAD (its tasks take integer indices, for_all_neighbor passes
structs: SURVEY B-14).  The semantics implemented here are the index-based reading pinned by the oracle
(oracle/sph_oracle_pbf.inc): fluid + boundary neighbours, kernels evaluated on the positions of the step's
start, update_all_pos as "move all particles, then XSPH" (csrc/sph_sweeps_pbf.cuh)."""
from . import _lib
from .fields import FetchedField
from .solver_base import solver_base


class pbf_solver(solver_base):
    solver_name = 'pbf'

    def __init__(self, particle_system, config):
        super(pbf_solver, self).__init__(particle_system, config)
        self.constrain = FetchedField(self.ps, _lib.F_SCALAR_A)                       # PBF:11-15
        self.pos_predict = FetchedField(self.ps, _lib.F_VEC_C, 3)
        self.delta_pos = FetchedField(self.ps, _lib.F_FORCE_B, 3)
        self.constrain_derivative = FetchedField(self.ps, _lib.F_FORCE_A, 3)
        self.pbf_lambda = FetchedField(self.ps, _lib.F_SCALAR_B)
        self.epsilon = 1.0e-6                                                         # PBF:16
        self.k = 1e-7                                                                 # PBF:18 tension
        self.c = 9e-6                                                                 # PBF:19 viscosity
        self.s_corr_factor = 0.3                                                      # PBF:20

    def externel_force_predict_pos(self):                                             # PBF:26-30
        self.ps.phase(_lib.PH_PBF_PREDICT)

    def compute_all_lambda(self):                                                     # PBF:32-52
        self.ps.phase(_lib.PH_PBF_LAMBDA)

    def compute_all_delta_pos(self):                                                  # PBF:55-65
        self.ps.phase(_lib.PH_PBF_DELTA_POS)

    def update_all_pos(self):                                                         # PBF:67-96
        self.ps.phase(_lib.PH_PBF_UPDATE_POS)

    def step(self):                                                                   # PBF:176-186
        self._full_step(1)
