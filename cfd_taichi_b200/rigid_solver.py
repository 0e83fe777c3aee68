"""rigid_solver -- drop-in mirror of the reference class (rigid_solver.py:4-234).

step() is one stream-ordered device kernel (sph_rigid_step): the force and torque reductions over the
rigid particles (RS:35-38, 121-123), the attitude / rotation update (RS:118-141), the wall contact scan
with its impulse (RS:53-94) and the particle / vertex / centroid move (RS:96-104).  The O(1) 3x3 algebra
runs in the same kernel, so a rigid step costs no host round trip.
"""
import torch

from . import _lib
from .fields import HostScalar


class rigid_solver:

    def __init__(self, particle_system, config):
        solid_config = config.get('solid')
        solver_config = config.get('solver')
        scene_config = config.get('scene')
        self.ps = particle_system
        self._lib = particle_system._lib
        self.delta_time = _RigidScalar(self, 'delta_time', solver_config.get('delta_time'))   # RS:12-13
        self.simulate_cnt = HostScalar(0)
        self.gravity = scene_config.get('gravity')
        self.rho = solid_config.get('rho_0')
        self.particle_count = self.ps.rigid_particles_num
        self.omega = _RigidVec(self, 'omega')                                          # RS:20-22
        self.attitude = _RigidVec(self, 'attitude')
        self.mass = _RigidScalar(self, 'mass', 0.0)
        self.v_decay_proportion = 0.1                                                  # RS:24
        self.run_once_flag = False

    def state(self):
        return self.ps.rigid_state()

    def compute_sum_mass(self):                                                        # RS:156-162
        print('rigid mass is {}'.format(self.state().mass))

    def run_once(self):                                                                # RS:212-214
        self.compute_sum_mass()

    def step(self):                                                                    # RS:216-234
        if not self.run_once_flag:
            self.run_once_flag = True
        self.simulate_cnt[None] += 1
        _lib.check(self._lib.sph_rigid_step(self.ps._h, self.ps._stream()), self.ps._h)

    def sync_fields(self):
        """Refresh the Python-visible centroid / inertia fields from the device state (the reference
        mutates ps.rigid_centriod and ps.rigid_inertia_tensor_inv in place, RS:104, 141)."""
        info = self.state()
        dev = self.ps._device
        self.ps.rigid_centriod.tensor.copy_(torch.tensor(list(info.centroid), dtype=torch.float32, device=dev))
        self.ps.rigid_inertia_tensor_inv.tensor.copy_(
            torch.tensor(list(info.inertia_inv), dtype=torch.float32, device=dev).reshape(3, 3))
        return info


class _RigidVec:
    def __init__(self, rs, name):
        self._rs, self._name = rs, name

    def __getitem__(self, idx):
        import numpy as np
        return np.array(list(getattr(self._rs.state(), self._name)), dtype=np.float32)


class _RigidScalar:
    def __init__(self, rs, name, default):
        self._rs, self._name, self._default = rs, name, default

    def __getitem__(self, idx):
        return float(getattr(self._rs.state(), self._name))
