"""ParticleSystem -- drop-in mirror of the reference class (ParticleSystem.py:28-507).

Same constructor (`ParticleSystem(config)`), same public attributes and methods that main.py and the
solvers touch; the state lives in torch CUDA tensors (float4 SoA) and every kernel is the
hand-written sm_100a library behind include/sph_b200.h.  There is no CPU path.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib, scene, slab as slab_plan
from .fields import DeviceScalar, FetchedField, HostScalar, ParticleFields, TensorField


class ParticleSystem:
    material_fluid = 0            # PS:74-76
    material_solid_boundary = 1
    material_solid = 2

    def __init__(self, config, device=None, strict=None, solver_name=None, ghost_capacity=0,
                 max_neighbors=None, base_dir=None, slab=None):
        """`slab=(rank, nranks[, cuts])`: this process holds one x-slab of the domain (multi-GPU);
        `cuts` overrides the histogram-balanced column cuts (nranks + 1 increasing x-cell columns);
        torch.distributed must be initialised (it carries the NCCL id; the exchange itself is in the library)."""
        if not torch.cuda.is_available():
            raise _lib.SphError("ParticleSystem needs a CUDA device: the B200 SPH path has no CPU fallback")
        self._lib = _lib.load()
        self.config = config
        scene_config = config.get('scene')
        solver_config = config.get('solver')
        fluid_config = config.get('fluid')
        solid_config = config.get('solid', {})
        self._device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self._strict = bool(solver_config.get('strict', False) if strict is None else strict)
        self._max_neighbors = int(max_neighbors or solver_config.get('max_neighbors', 0))
        self._ghost_capacity = int(ghost_capacity)
        self._base_dir = base_dir or os.getcwd()
        self._h = None
        self._solver_name = None

        self.exist_rigid = HostScalar(1 if solid_config else 0)                     # PS:39-40
        self.active_rigid = HostScalar(1 if solid_config.get('active', False) else 0)  # PS:63-64

        self.water_size = list(fluid_config.get('water_size'))                      # PS:78-83
        self.start_pos = list(fluid_config.get('start_pos'))
        self.particle_radius = scene_config.get('particle_radius')
        self.particle_diameter = self.particle_radius * 2
        self.support_radius = 4 * self.particle_radius
        self.particle_m = 1000 * (self.particle_radius ** 3) * 8
        self.box_max = list(scene_config.get('box_max'))
        self.box_min = list(scene_config.get('box_min'))

        self.particle_num, self.boundary_particles_num, grid_num = scene.derive_sizes(config)
        self.grid_num = list(grid_num)                                              # PS:100-102
        self._3d_to_1d_tran = [1, self.grid_num[0] * self.grid_num[2], self.grid_num[0]]
        print('Boundary particle count: {}k'.format(self.boundary_particles_num / 1000))

        self._slab = None
        self._owned_ids = None
        self._n_owned_cap = self.particle_num
        if slab is not None and slab[1] > 1:
            rank, nranks = int(slab[0]), int(slab[1])
            hist, _ = slab_plan.column_histogram(config)
            cuts = list(slab[2]) if len(slab) > 2 and slab[2] is not None else slab_plan.plan_cuts(hist, nranks)
            if len(cuts) != nranks + 1 or cuts[0] != 0 or cuts[-1] != len(hist) or any(b <= a for a, b in zip(cuts, cuts[1:])):
                raise ValueError("slab cuts must be %d increasing x-cell columns from 0 to %d: %r" % (nranks + 1, len(hist), cuts))
            ids, _, _ = slab_plan.owned_lattice_ids(config, cuts[rank], cuts[rank + 1])
            owned0, owned_cap, ghost_cap = slab_plan.capacities(config, cuts, rank)
            assert owned0 == len(ids)
            self._slab = dict(rank=rank, nranks=nranks, cuts=cuts, col_lo=cuts[rank], col_hi=cuts[rank + 1])
            self._owned_ids = ids
            self._n_owned_cap = owned_cap
            self._ghost_capacity = ghost_cap

        dev = self._device
        n, nb = self._n_owned_cap, self.boundary_particles_num
        ncap = n + self._ghost_capacity
        # caller-owned state, float4 SoA (see include/sph_b200.h enum SphField)
        self._pos4 = torch.zeros((ncap, 4), dtype=torch.float32, device=dev)
        self._vel4 = torch.zeros((ncap, 4), dtype=torch.float32, device=dev)
        self._acc4 = torch.zeros((ncap, 4), dtype=torch.float32, device=dev)
        self._bpos4 = torch.zeros((max(nb, 1), 4), dtype=torch.float32, device=dev)
        self._gid = torch.zeros((ncap,), dtype=torch.int32, device=dev) if self._slab else None
        self._rgb = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        self.rgba = TensorField(torch.tensor([0.0, 0.26, 0.68, 1.0], device=dev).repeat(n, 1))   # PS:113,152
        self.rgb = TensorField(torch.tensor([0.0, 0.28, 1.0], device=dev).repeat(n, 1))          # PS:116-117
        self._rgb.copy_(torch.tensor([0.0, 0.28, 1.0], device=dev).expand(n, 3))                 # PS:227

        # rigid body (PS:41-64)
        self.rigid_particles_num = 0
        self.rigid_vertex_count = 0
        self.rigid_centriod = TensorField(torch.zeros(3, dtype=torch.float32, device=dev))
        self._rigid_points = None
        if self.exist_rigid[None] == 1:
            pts, verts, faces = scene.rigid_points_from_config(solid_config, self._base_dir)
            self._rigid_points, self._rigid_faces = pts, faces
            self._rigid_vertices_local = np.ascontiguousarray(verts, dtype=np.float32)   # before rotation / offset
            self.voxel_radius = solid_config.get('voxel_radius')
            self.rigid_pos_offset = solid_config.get('pos_offset')
            self.rigid_attitude_offset = [a / 180.0 * math.pi for a in solid_config.get('attitude_offset')]
            self.rigid_rho = solid_config.get('rho_0')
            self.rigid_vertex_count = verts.shape[0]
            self.rigid_particles_num = pts.shape[0]
            self._rverts4 = torch.zeros((max(verts.shape[0], 1), 4), dtype=torch.float32, device=dev)
            self._rverts4[:verts.shape[0], :3] = torch.from_numpy(verts.astype(np.float32)).to(dev)
            self._rigid_vertices = self._rverts4[:verts.shape[0], :3]
            self.rigid_vertices = TensorField(self._rigid_vertices)
            self.rigid_inertia_tensor = TensorField(torch.zeros((3, 3), dtype=torch.float32, device=dev))
            self.rigid_inertia_tensor_inv = TensorField(torch.zeros((3, 3), dtype=torch.float32, device=dev))
        nr = self.rigid_particles_num
        self._rpos4 = torch.zeros((max(nr, 1), 4), dtype=torch.float32, device=dev)
        self._rvel4 = torch.zeros((max(nr, 1), 4), dtype=torch.float32, device=dev)
        self._rforce4 = torch.zeros((max(nr, 1), 4), dtype=torch.float32, device=dev)
        self._rkin = {k: torch.zeros((max(nr, 1), 3), dtype=torch.float32, device=dev)
                      for k in ('acc', 'omega', 'alpha')}
        self._rrgb = torch.zeros((max(nr, 1), 3), dtype=torch.float32, device=dev)

        self.fluid_particles = ParticleFields(
            pos=TensorField(self._pos4[:n, :3]), vel=TensorField(self._vel4[:n, :3]),
            acc=TensorField(self._acc4[:n, :3]), rgb=TensorField(self._rgb),
            index=TensorField(torch.arange(n, dtype=torch.int32, device=dev)),
            belong_grid=_BelongGrid(self, 'fluid'))
        self.boundary_particles = ParticleFields(
            pos=TensorField(self._bpos4[:nb, :3]), volume=TensorField(self._bpos4[:nb, 3]),
            index=TensorField(torch.arange(nb, dtype=torch.int32, device=dev)))
        self.rigid_particles = ParticleFields(
            pos=TensorField(self._rpos4[:nr, :3]), volume=TensorField(self._rpos4[:nr, 3]),
            vel=TensorField(self._rvel4[:nr, :3]), mass=TensorField(self._rvel4[:nr, 3]),
            force=TensorField(self._rforce4[:nr, :3]), acc=TensorField(self._rkin['acc'][:nr]),
            omega=TensorField(self._rkin['omega'][:nr]), alpha=TensorField(self._rkin['alpha'][:nr]),
            rgb=TensorField(self._rrgb[:nr]),
            index=TensorField(torch.arange(nr, dtype=torch.int32, device=dev)))

        # adaptive dt written by DFSPH (DF:119), read by the rigid solver (RS:223-224)
        self.delta_time = DeviceScalar(lambda: self.read_stats().ps_delta_time, lambda v: None)

        self.init_particle_pos()                                                    # PS:119
        if self.exist_rigid[None] == 1:
            self.init_rigid_particles_pos()                                         # PS:120-121
        self._create_handle(solver_name or solver_config.get('name'))
        self.init_particles_data()                                                  # PS:122

        print('Fluid particle count: {}k'.format(self.particle_num / 1000))
        print('Solid particle count: {}k'.format(self.rigid_particles_num / 1000))
        print('Particle mass: {}'.format(self.particle_m))
        print('Grid: {}, Grid count: {}'.format(self.grid_num, self.grid_num[0] * self.grid_num[1] * self.grid_num[2]))

    # ------------------------------------------------------------------------------------------
    # handle management
    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def _create_handle(self, solver_name):
        if solver_name not in _lib.SOLVER_IDS:
            raise _lib.SphError("solver '%s' has no CUDA path in this build" % solver_name)
        if self._h is not None:
            _lib.check(self._lib.sph_destroy(self._h))
            self._h = None
        scene_config, solver_config = self.config['scene'], self.config['solver']
        solid_config = self.config.get('solid', {})
        cfg = _lib.SphConfig()
        for k in range(3):
            cfg.box_min[k] = self.box_min[k]
            cfg.box_max[k] = self.box_max[k]
            cfg.grid_num[k] = self.grid_num[k]
        cfg.particle_radius = self.particle_radius
        cfg.gravity = scene_config.get('gravity')
        cfg.delta_time = solver_config.get('delta_time')
        cfg.boundary_handle = 1 if solver_config.get('boundary_handle', True) else 0   # SB:31
        cfg.fs_couple = 1 if solver_config.get('fs_couple', True) else 0               # SB:32
        cfg.solver = _lib.SOLVER_IDS[solver_name]
        cfg.n_fluid = self._n_owned_cap
        cfg.n_boundary = self.boundary_particles_num
        cfg.n_rigid = self.rigid_particles_num
        cfg.active_rigid = int(self.active_rigid[None])
        cfg.max_neighbors = self._max_neighbors
        cfg.strict = 1 if self._strict else 0
        cfg.n_ghost_capacity = self._ghost_capacity
        cfg.rigid_rho = solid_config.get('rho_0', 0.0) if solid_config else 0.0
        cfg.use_graph = 0   # reserved (include/sph_b200.h)
        h = ctypes.c_void_p()
        with torch.cuda.device(self._device):
            rc = self._lib.sph_create(ctypes.byref(cfg), self._device.index, ctypes.byref(h))
        if rc != 0:
            msg = self._lib.sph_last_error(h if h else None)
            if h:
                self._lib.sph_destroy(h)
            raise _lib.SphError("sph_create failed (%d): %s" % (rc, msg.decode()))
        self._h = h
        self._solver_name = solver_name
        self._bind_all()
        if self._slab:
            self._join_slab()

    def _bind_all(self):
        L, h = self._lib, self._h
        _lib.check(L.sph_bind(h, _lib.F_FLUID_POS, self._pos4.data_ptr(), self._pos4.shape[0]), h)
        _lib.check(L.sph_bind(h, _lib.F_FLUID_VEL, self._vel4.data_ptr(), self._vel4.shape[0]), h)
        _lib.check(L.sph_bind(h, _lib.F_FLUID_ACC, self._acc4.data_ptr(), self._acc4.shape[0]), h)
        if self.boundary_particles_num > 0:
            _lib.check(L.sph_bind(h, _lib.F_BOUNDARY_POS, self._bpos4.data_ptr(), self._bpos4.shape[0]), h)
        if self.rigid_particles_num > 0:
            _lib.check(L.sph_bind(h, _lib.F_RIGID_POS, self._rpos4.data_ptr(), self._rpos4.shape[0]), h)
            _lib.check(L.sph_bind(h, _lib.F_RIGID_VEL, self._rvel4.data_ptr(), self._rvel4.shape[0]), h)
            _lib.check(L.sph_bind(h, _lib.F_RIGID_FORCE, self._rforce4.data_ptr(), self._rforce4.shape[0]), h)
            _lib.check(L.sph_bind(h, _lib.F_RIGID_VERTICES, self._rverts4.data_ptr(), self.rigid_vertex_count), h)

    def _join_slab(self):
        """Bind the global ids, set the owned count and join the NCCL communicator (sph_comm_init)."""
        import torch.distributed as dist
        L, h, sl = self._lib, self._h, self._slab
        n0 = len(self._owned_ids)
        self._gid[:n0] = torch.from_numpy(self._owned_ids.astype(np.int32)).to(self._device)
        _lib.check(L.sph_bind(h, _lib.F_FLUID_GID, self._gid.data_ptr(), self._gid.shape[0]), h)
        _lib.check(L.sph_set_counts(h, n0, 0), h)
        buf = ctypes.create_string_buffer(128)
        if sl['rank'] == 0:
            _lib.check(L.sph_comm_unique_id(buf))
        t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=self._device)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().tolist())
        _lib.check(L.sph_comm_init(h, raw, sl['rank'], sl['nranks'], sl['col_lo'], sl['col_hi']), h)

    def comm_info(self):
        out = (ctypes.c_int32 * 8)()
        _lib.check(self._lib.sph_comm_info(self._h, out), self._h)
        return dict(owned=out[0], ghosts=out[1], sent=(out[2], out[3]), received=(out[4], out[5]), rank=out[6],
                    nranks=out[7])

    def local_count(self):
        """Particles this handle currently sorts (owned + ghosts); == particle_num on one GPU."""
        if not self._slab:
            return self.particle_num
        i = self.comm_info()
        return i['owned'] + i['ghosts']

    def owned_state(self):
        """(gid, pos, vel) of the particles this rank owns, as numpy arrays (tests / output writers)."""
        n = self.comm_info()['owned'] if self._slab else self.particle_num
        gid = self._gid[:n].cpu().numpy() if self._slab else np.arange(n, dtype=np.int32)
        return gid, self._pos4[:n, :3].cpu().numpy(), self._vel4[:n].cpu().numpy()

    def _ensure_solver(self, solver_name):
        """Solvers are located by name (main.py:65-68); a solver class built on a ParticleSystem whose
        config names another solver re-creates the handle with its own constants (SB:24-26 vs WC:18-20)."""
        if solver_name != self._solver_name:
            self._create_handle(solver_name)
            _lib.check(self._lib.sph_init_boundary(self._h, self._stream()), self._h)
            if self.exist_rigid[None]:
                _lib.check(self._lib.sph_init_rigid(self._h, self._stream()), self._h)

    def close(self):
        if getattr(self, '_h', None) is not None:
            self._lib.sph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _fetch(self, field_id, width, dtype, count=None):
        n = self.local_count() if count is None else count
        out = torch.empty((n, width), dtype=dtype, device=self._device)
        _lib.check(self._lib.sph_fetch(self._h, field_id, out.data_ptr(), n, self._stream()), self._h)
        return out

    def read_stats(self):
        st = _lib.SphStats()
        _lib.check(self._lib.sph_read_stats(self._h, ctypes.byref(st)), self._h)
        return st

    def phase(self, phase_id):
        _lib.check(self._lib.sph_phase(self._h, phase_id, self._stream()), self._h)

    # ------------------------------------------------------------------------------------------
    # reference methods
    # ------------------------------------------------------------------------------------------
    def compute_boundary_particles_count(self):                                     # PS:129-137
        return scene.derive_sizes(self.config)[1]

    def init_particle_pos(self):                                                    # PS:139-195
        """Fluid lattice and boundary shell generated on the device (sph_init_fluid_lattice /
        sph_init_boundary_shell); scene.init_*_positions is the numpy statement of the same formulas
        that the tests compare against."""
        n, nb = self.particle_num, self.boundary_particles_num
        sc, fl = self.config['scene'], self.config['fluid']
        lat = _lib.SphLattice()
        lat.particle_radius = sc['particle_radius']
        for k in range(3):
            lat.start_pos[k], lat.water_size[k] = fl['start_pos'][k], fl['water_size'][k]
            lat.box_min[k], lat.box_max[k] = sc['box_min'][k], sc['box_max'][k]
        L = _lib.load()
        dev_index = self._device.index if self._device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(self._device).cuda_stream
        ids, n_local = None, n
        if self._owned_ids is not None:
            ids = torch.from_numpy(np.ascontiguousarray(self._owned_ids, dtype=np.int32)).to(self._device)
            n_local = ids.shape[0]
        _lib.check(L.sph_init_fluid_lattice(ctypes.byref(lat), n, ids.data_ptr() if ids is not None else None, n_local,
                                            self._pos4.data_ptr(), dev_index, stream))
        if nb > 0:
            _lib.check(L.sph_init_boundary_shell(ctypes.byref(lat), nb, self._bpos4.data_ptr(), dev_index, stream))
        torch.cuda.current_stream(self._device).synchronize()   # ids may be freed after this call

    def init_rigid_particles_pos(self):                                             # PS:198-223
        a = self.rigid_attitude_offset
        R = torch.from_numpy(scene.rotation3d(a[0], a[2], a[1])).to(self._device)
        off = torch.tensor(self.rigid_pos_offset, dtype=torch.float32, device=self._device)
        nr = self.rigid_particles_num
        p = torch.from_numpy(self._rigid_points).to(self._device)
        self._rpos4[:nr, :3] = _rot_rows(R, p) + off
        self._rverts4[:self.rigid_vertex_count, :3] = _rot_rows(R, self._rigid_vertices.clone()) + off

    def init_particles_data(self):                                                  # PS:225-247
        self.reset_boundary_grids()
        self.update_boundary_grids()
        self.reset_grid()
        self.update_grid()
        self.compute_all_boundary_volume()
        if self.exist_rigid[None]:
            self.init_rigid_particles_data()

    def reset_boundary_grids(self):                                                 # PS:322-327
        pass  # the CSR rebuild in sph_init_boundary replaces deactivate()

    def update_boundary_grids(self):                                                # PS:329-335
        self._boundary_dirty = True

    def compute_all_boundary_volume(self):                                          # PS:309-320
        # boundary grid + Akinci volumes are one library call; volumes land in boundary_particles.volume
        _lib.check(self._lib.sph_init_boundary(self._h, self._stream()), self._h)
        self._boundary_dirty = False

    def init_rigid_particles_data(self):                                            # PS:249-295
        _lib.check(self._lib.sph_init_rigid(self._h, self._stream()), self._h)
        info = self.rigid_state()
        dev = self._device
        self.rigid_centriod.tensor.copy_(torch.tensor(list(info.centroid), dtype=torch.float32, device=dev))
        self.rigid_inertia_tensor.tensor.copy_(torch.tensor(list(info.inertia), dtype=torch.float32, device=dev).reshape(3, 3))
        self.rigid_inertia_tensor_inv.tensor.copy_(
            torch.tensor(list(info.inertia_inv), dtype=torch.float32, device=dev).reshape(3, 3))
        self._rrgb[:self.rigid_particles_num] = torch.tensor([1.0, 0.0, 0.0], device=dev)            # PS:295
        print("Centroid: {}".format(list(info.centroid)))
        print("Intertia tensor: {}".format([list(info.inertia)[k * 3:k * 3 + 3] for k in range(3)]))

    def rigid_state(self):
        """Device-resident rigid-body state (centroid, inertia, velocities ...); synchronises."""
        info = _lib.SphRigidInfo()
        _lib.check(self._lib.sph_rigid_state(self._h, ctypes.byref(info)), self._h)
        return info

    def reset_grid(self):                                                           # PS:368-373
        pass  # cell counters are cleared inside the grid build (sph_phase BUILD_GRID)

    def update_grid(self):                                                          # PS:382-386
        if getattr(self, '_boundary_dirty', True) and self.boundary_particles_num > 0:
            _lib.check(self._lib.sph_init_boundary(self._h, self._stream()), self._h)
            self._boundary_dirty = False
        self.phase(_lib.PH_BUILD_GRID)

    def test(self):                                                                 # PS:376-379
        self.reset_grid()
        self.update_grid()
        self.check_all_grid()

    def check_all_grid(self):                                                       # PS:471-484
        start = self._fetch_raw(_lib.F_CELL_START, self.grid_count + 1)
        total = int(start[-1].item())
        print('Check pass!' if total == self.particle_num else 'Fail!')
        return total == self.particle_num

    @property
    def grid_count(self):
        return self.grid_num[0] * self.grid_num[1] * self.grid_num[2]

    def _fetch_raw(self, field_id, count):
        out = torch.empty((count,), dtype=torch.int32, device=self._device)
        _lib.check(self._lib.sph_fetch(self._h, field_id, out.data_ptr(), count, self._stream()), self._h)
        return out

    def cell_indices_1d(self):
        """1-D cell id per fluid particle, original order (PS:486-494)."""
        return self._fetch_raw(_lib.F_CELL1D, self.local_count())

    def cell_start(self):
        return self._fetch_raw(_lib.F_CELL_START, self.grid_count + 1)

    def sorted_index(self):
        return self._fetch_raw(_lib.F_SORTED_INDEX, self.local_count())

    def neighbour_counts(self):
        """get_neighbour_count(i) for every fluid particle (PS:424-445); needs the step's lists."""
        return self._fetch(_lib.F_NEIGHBOR_COUNT, 1, torch.int32).reshape(-1)

    def get_max_neighbor_particle_index(self):                                      # PS:409-422
        cnt = self.neighbour_counts().cpu().numpy()
        max_index = slab_plan.racy_argmax(cnt)      # one-thread semantics of the racy arg-max
        print('max_index is {}, length is {}'.format(max_index, int(cnt.max()) if cnt.size else -1))
        return max_index

    def update_mesh_vextics(self):                                                  # PS:298-299
        self.mesh_vertices = self._rigid_vertices.cpu().numpy()


class _BelongGrid:
    """fluid_particles.belong_grid (PS:13,397): 3-D cell of each particle, derived from the 1-D id."""

    def __init__(self, ps, which):
        self._ps = ps

    def to_numpy(self):
        ps = self._ps
        c1 = ps.cell_indices_1d().cpu().numpy().astype(np.int64)
        gx, gxz = ps.grid_num[0], ps.grid_num[0] * ps.grid_num[2]
        y = c1 // gxz
        rem = c1 - y * gxz
        z = rem // gx
        x = rem - z * gx
        return np.stack([x, y, z], axis=1).astype(np.int32)


def _rot_rows(R, p):
    """Row-wise R @ p with the reference's left-to-right f32 accumulation (no fused reordering)."""
    return torch.stack([(R[k, 0] * p[:, 0] + R[k, 1] * p[:, 1]) + R[k, 2] * p[:, 2] for k in range(3)], dim=1)
