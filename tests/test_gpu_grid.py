"""GPU parity: neighbour search (cell hash, prefix sum, counting sort, neighbour counts, Akinci
boundary volumes) against the CPU oracle.  Integer outputs are compared bit-exactly."""
import numpy as np
import pytest
import torch

from cfd_taichi_b200 import scenes
from conftest import quiet_ps
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _pair(cfg, solver="dfsph", strict=True, **kw):
    ps = quiet_ps(cfg, strict=strict, solver_name=solver, **kw)
    o = O.Oracle(cfg, solver=solver, threads=8)
    return ps, o


@pytest.mark.parametrize("name", ["small_block", "breaking_dam_30k", "default"])
@pytest.mark.parametrize("strict", [True, False])
def test_grid_bit_exact_on_lattice(built, name, strict):
    cfg = scenes.shipped(name, "dfsph")
    ps, o = _pair(cfg, strict=strict)
    assert np.array_equal(ps.cell_indices_1d().cpu().numpy(), o.field("cell1"))
    assert np.array_equal(ps.cell_start().cpu().numpy(), o.field("cell_start"))
    assert np.array_equal(ps.sorted_index().cpu().numpy(), o.field("cell_items"))
    assert np.array_equal(ps.fluid_particles.belong_grid.to_numpy(), o.field("cell3"))
    assert ps.check_all_grid()
    if cfg["solver"].get("boundary_handle", True):
        bv = ps.boundary_particles.volume.to_numpy()
        if strict:
            assert np.array_equal(bv, o.field("bvol"))
        else:
            assert np.allclose(bv, o.field("bvol"), rtol=1e-5, atol=0)
    ps.close(); o.close()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_grid_and_counts_bit_exact_on_jittered_particles(built, seed):
    # randomised property test: jitter U(-0.2 d, 0.2 d), integer outputs still identical
    cfg = scenes.shipped("small_block", "dfsph")
    ps, o = _pair(cfg, strict=False)
    rng = np.random.default_rng(seed)
    pos = o.field("pos")
    pos += rng.uniform(-0.01, 0.01, size=pos.shape).astype(np.float32)
    ps.fluid_particles.pos.from_numpy(pos)
    ps.update_grid()
    o.phase("reset_grid_update_grid")
    c1 = ps.cell_indices_1d().cpu().numpy()
    assert np.array_equal(c1, o.field("cell1"))
    start, order = ps.cell_start().cpu().numpy(), ps.sorted_index().cpu().numpy()
    assert np.array_equal(start, o.field("cell_start"))
    assert np.array_equal(order, o.field("cell_items"))
    # properties: permutation, cell-contiguous, stable tie-break by original index
    assert np.array_equal(np.sort(order), np.arange(ps.particle_num))
    cs = c1[order]
    assert np.all(np.diff(cs) >= 0) and np.all(np.diff(order)[np.diff(cs) == 0] > 0)
    from cfd_taichi_b200 import _lib
    ps.phase(_lib.PH_DF_INITIALIZE)
    o.phase("neighbour_counts")
    assert np.array_equal(ps.neighbour_counts().cpu().numpy(), o.field("nbr_count"))
    ps.close(); o.close()


def test_ragged_and_empty_cells(built):
    # a thin sheet of fluid: most cells empty, a few very full ones after squeezing particles together
    cfg = scenes.make_scene([1.5, 1.5, 1.5], [0.3, 0.3, 0.3], [0.5, 0.1, 0.5], "dfsph", 1e-3)
    ps, o = _pair(cfg, strict=True)
    pos = o.field("pos")
    pos[:, 0] = 0.3 + (pos[:, 0] - 0.3) * 0.25        # 4x compression along x: up to 64 particles per cell
    ps.fluid_particles.pos.from_numpy(pos)
    ps.update_grid()
    o.phase("reset_grid_update_grid")
    assert np.array_equal(ps.cell_start().cpu().numpy(), o.field("cell_start"))
    assert np.array_equal(ps.sorted_index().cpu().numpy(), o.field("cell_items"))
    counts = np.diff(o.field("cell_start"))
    assert counts.max() >= 16 and (counts == 0).mean() > 0.9
    from cfd_taichi_b200 import _lib
    ps.phase(_lib.PH_DF_INITIALIZE)
    o.phase("neighbour_counts")
    assert np.array_equal(ps.neighbour_counts().cpu().numpy(), o.field("nbr_count"))
    st = ps.read_stats()
    assert st.error_flags == 0 and st.max_neighbors_seen == o.field("nbr_count").max()
    ps.close(); o.close()


def test_particle_outside_grid_is_flagged_not_fatal(built):
    cfg = scenes.shipped("small_block", "dfsph")
    ps = quiet_ps(cfg, strict=True)
    pos = ps.fluid_particles.pos.to_numpy()
    pos[5] = [-1.0, 0.5, 0.5]                          # the reference only prints here (PS:393-395)
    ps.fluid_particles.pos.from_numpy(pos)
    ps.update_grid()
    assert ps.read_stats().error_flags & 1
    assert ps.check_all_grid()                         # still addressable: clamped into the grid
    ps.close()


def test_list_overflow_is_flagged(built):
    cfg = scenes.shipped("small_block", "dfsph")
    ps = quiet_ps(cfg, strict=True, max_neighbors=16)
    from cfd_taichi_b200 import _lib
    ps.phase(_lib.PH_DF_INITIALIZE)
    assert ps.read_stats().error_flags & 2
    ps.close()


def test_full_size_grid_properties_1m(built):
    # BASELINE size (1 M particles): size-independent properties instead of an oracle run
    cfg = scenes.breaking_dam(100)
    ps = quiet_ps(cfg, strict=False)
    n = ps.particle_num
    c1 = ps.cell_indices_1d()
    start, order = ps.cell_start(), ps.sorted_index()
    assert int(start[-1]) == n and int(start[0]) == 0
    assert bool((start[1:] >= start[:-1]).all())
    assert torch.equal(torch.sort(order).values.cpu(), torch.arange(n, dtype=torch.int32))
    cs = c1[order.long()]
    d = cs[1:] - cs[:-1]
    assert bool((d >= 0).all())
    same = d == 0
    assert bool(((order[1:] - order[:-1])[same] > 0).all())
    # histogram of cell ids == differences of the prefix sum (a checksum of checksums)
    hist = torch.bincount(c1.long(), minlength=ps.grid_count)
    assert torch.equal(hist.to(torch.int32), (start[1:] - start[:-1]))
    # cell ids follow floor(pos / 0.1f) exactly
    p = ps.fluid_particles.pos.tensor
    c3 = torch.floor(p / torch.tensor(0.1, dtype=torch.float32, device=p.device)).to(torch.int64)
    gx, gz = ps.grid_num[0], ps.grid_num[2]
    assert torch.equal(c3[:, 0] + gx * gz * c3[:, 1] + gx * c3[:, 2], c1.long())
    ps.close()


@pytest.mark.parametrize("name,solver", [("small_block", "dfsph"), ("breaking_dam_30k", "wcsph"), ("dam_flush_cube", "dfsph")])
def test_device_side_init_equals_numpy_statement(built, name, solver):
    # init_particle_pos (PS:139-195) on the device vs the numpy statement of the same formulas
    from cfd_taichi_b200 import scene
    cfg = scenes.shipped(name, solver)
    cfg.pop("solid", None)            # the mesh asset of dam_flush_cube is not part of this repository
    ps = quiet_ps(cfg, strict=True)
    n, nb = ps.particle_num, ps.boundary_particles_num
    assert np.array_equal(ps._pos4[:n, :3].cpu().numpy(), scene.init_fluid_positions(cfg, n))
    assert np.array_equal(ps.boundary_particles.pos.to_numpy(), scene.init_boundary_positions(cfg, nb))
    ps.close()


def test_device_side_init_integer_lattice_beyond_2_24(built):
    # 256^3 = 2^24 particles: the reference's f32 index arithmetic stops being exact, integer lattice instead
    import ctypes
    from cfd_taichi_b200 import _lib, scene
    cfg = scenes.breaking_dam(256)
    n = scene.derive_sizes(cfg)[0]
    assert n == 1 << 24
    lat = _lib.SphLattice()
    lat.particle_radius = cfg["scene"]["particle_radius"]
    for k in range(3):
        lat.start_pos[k], lat.water_size[k] = cfg["fluid"]["start_pos"][k], cfg["fluid"]["water_size"][k]
        lat.box_min[k], lat.box_max[k] = cfg["scene"]["box_min"][k], cfg["scene"]["box_max"][k]
    ids = torch.tensor([0, 255, 256, 65535, 65536, (1 << 24) - 1, 12345678], dtype=torch.int32, device="cuda")
    out = torch.zeros((ids.shape[0], 4), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().sph_init_fluid_lattice(ctypes.byref(lat), n, ids.data_ptr(), ids.shape[0], out.data_ptr(), 0,
                                                  torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    i = ids.cpu().numpy().astype(np.int64)
    want = np.stack([i % 256, i // 65536, (i // 256) % 256], axis=1).astype(np.float32) * np.float32(0.025) * np.float32(2) + np.float32(0.1)
    assert np.array_equal(out[:, :3].cpu().numpy(), want)
