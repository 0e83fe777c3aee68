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
        if ps._slab is not None:
            self._pre_compute_slab()
        else:
            max_index = ps.get_max_neighbor_particle_index()
            if max_index >= 0:
                _lib.check(self._lib.sph_pcisph_delta(ps._h, int(max_index), ps._stream()), ps._h)
        st = self.stats()
        self.delta[None] = st.pc_delta
        print('PCISPH parameter delta: {}, beta: {}'.format(self.delta[None], self.beta))

    def _pre_compute_slab(self):
        """x-slabs: the arg-max particle of the GLOBAL domain (PS:409-422 over all ranks' counts, by global
        particle id) decides; its owner computes delta and every rank takes that value."""
        import torch
        import torch.distributed as dist
        ps = self.ps
        world = dist.get_world_size()
        n = ps.comm_info()['owned']
        cnt = ps.neighbour_counts()[:n].to(torch.int32)
        gid = ps._gid[:n]
        sizes = [torch.zeros(1, dtype=torch.int64, device=cnt.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device=cnt.device))
        cap = int(max(int(s.item()) for s in sizes))
        pad = torch.full((2, cap), -1, dtype=torch.int32, device=cnt.device)
        pad[0, :n], pad[1, :n] = gid, cnt
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        from . import slab as slab_plan
        max_index, src = slab_plan.global_argmax(
            ps.particle_num, [(part[0, :int(sz.item())].cpu().numpy(), part[1, :int(sz.item())].cpu().numpy())
                              for part, sz in zip(parts, sizes)])
        delta = torch.zeros(1, dtype=torch.float32, device=cnt.device)
        if max_index >= 0:
            if src == dist.get_rank():
                local = int(torch.nonzero(gid == max_index)[0].item())
                _lib.check(self._lib.sph_pcisph_delta(ps._h, local, ps._stream()), ps._h)
                delta[0] = self.stats().pc_delta
            dist.broadcast(delta, src=src)
            _lib.check(self._lib.sph_pcisph_set_delta(ps._h, ctypes.c_float(float(delta.item())), max_index, ps._stream()), ps._h)

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
