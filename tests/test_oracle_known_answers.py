"""CPU tests that pin the oracle (oracle/sph_oracle.c) with hand-derived known answers.

The reference ships no tests or golden vectors and taichi cannot be imported here, so these
known-answer checks plus the committed regression fixtures are the oracle's pin.
"""
import math

import numpy as np
import pytest

from cfd_taichi_b200 import scenes
from oracle import oracle as O

# (particle_num, boundary_particles_num, grid_num)
APPENDIX_C = {
    "default": (132479, 67202, (71, 71, 26)),
    "breaking_dam_30k": (29120, 21602, (51, 31, 16)),
    "dam_flush_cube": (56447, 21602, (51, 31, 16)),
    "small_block": (5879, 9002, (16, 31, 16)),
}


@pytest.mark.parametrize("name", sorted(APPENDIX_C))
def test_derived_sizes_shipped(name):
    assert O.derived_sizes(scenes.shipped(name)) == APPENDIX_C[name]


@pytest.mark.parametrize("n_side,gpus,expect", [
    (100, 1, (1000000, 191682, (151, 81, 53))),
    (160, 1, (4096000, 465122, (241, 121, 83))),
    (200, 1, (8000000, 725402, (301, 151, 103))),
    (200, 8, (64000000, 4950602, (2401, 151, 103))),
])
def test_derived_sizes_synthetic(n_side, gpus, expect):
    assert O.derived_sizes(scenes.breaking_dam(n_side, gpus_x=gpus)) == expect


def test_cubic_kernel_values():
    h = np.float32(0.1)
    # W(0) = 8 / (pi h^3)
    assert abs(O.cubic_kernel(0.0, h) - 8.0 / (math.pi * 1e-3)) < 1e-3 * 2546.479
    assert O.cubic_kernel(0.1000001, h) == 0.0
    # continuity at q = 0.5 and the two branch formulas
    k = 8.0 / (math.pi * 1e-3)
    assert abs(O.cubic_kernel(0.05, h) - k * 0.25) < 1e-2
    assert abs(O.cubic_kernel(0.075, h) - 2 * k * 0.25 ** 3) < 1e-2


def test_cubic_kernel_normalisation():
    # integral of W over the support ~ 1 (fine radial quadrature)
    h = np.float32(0.1)
    r = (np.arange(4000) + 0.5) * (0.1 / 4000)
    w = np.array([O.cubic_kernel(x, h) for x in r])
    assert abs(np.sum(4 * math.pi * r * r * w) * (0.1 / 4000) - 1.0) < 2e-3


def test_cubic_kernel_derivative_factor_six():
    # the reference's gradient is 6x the textbook one (SB:95-100): compare with 6 * dW/dr by finite differences
    h = np.float32(0.1)
    for r in (0.02, 0.04, 0.06, 0.09):
        g = O.cubic_kernel_derivative([r, 0.0, 0.0], h)
        fd = (O.cubic_kernel(r + 1e-4, h) - O.cubic_kernel(r - 1e-4, h)) / 2e-4
        assert abs(g[0] - 6.0 * fd) < 2e-2 * abs(6.0 * fd)
        assert g[1] == 0.0 and g[2] == 0.0
    assert np.all(O.cubic_kernel_derivative([0.0, 0.0, 0.0], h) == 0.0)          # q <= 1e-5
    assert np.all(O.cubic_kernel_derivative([0.2, 0.0, 0.0], h) == 0.0)          # q > 1


def test_cull_threshold_is_exact():
    # sqrt(r2) > h  <=>  r2 > T for every float around h*h
    h = np.float32(0.1)
    T = np.float32(O.cull_threshold(h))
    x = np.float32(h * h)
    vals = [x]
    for _ in range(40):
        vals.append(np.nextafter(vals[-1], np.float32(1)))
    y = x
    for _ in range(40):
        y = np.nextafter(y, np.float32(0))
        vals.append(y)
    for v in vals:
        assert (np.sqrt(np.float32(v)) > h) == (np.float32(v) > T)


@pytest.fixture(scope="module")
def small_dfsph():
    o = O.Oracle(scenes.shipped("small_block", "dfsph"), threads=4)
    yield o
    o.close()


def test_lattice_positions(small_dfsph):
    o = small_dfsph
    pos = o.field("pos")
    assert pos.shape == (5879, 3)
    # PS:150: fl(fl(k * 0.025f) * 2) + start ; 14 x 30 x 14 lattice minus the last site (SURVEY B-1)
    assert np.allclose(pos[0], [0.3, 0.5, 0.3])
    assert np.allclose(pos[-1], [0.3 + 12 * 0.05, 0.5 + 29 * 0.05, 0.3 + 13 * 0.05])
    d = pos[1] - pos[0]
    assert abs(d[0] - 0.05) < 1e-6 and d[1] == 0 and d[2] == 0


def test_grid_is_a_permutation_and_canonical(small_dfsph):
    o = small_dfsph
    start, items, cell1 = o.field("cell_start"), o.field("cell_items"), o.field("cell1")
    n = int(o.scalar("particle_num"))
    assert start[-1] == n                                       # check_all_grid (PS:471-484)
    assert np.array_equal(np.sort(items), np.arange(n))         # permutation
    cells_sorted = cell1[items]
    assert np.all(np.diff(cells_sorted) >= 0)                   # cell-contiguous
    same = np.diff(cells_sorted) == 0
    assert np.all(np.diff(items)[same] > 0)                     # ascending index inside a cell
    # 1-D id = x + gx*gz*y + gx*z (PS:102), cell3 = floor(pos / 0.1f)
    c3 = np.floor(o.field("pos") / np.float32(0.1)).astype(np.int64)
    assert np.array_equal(c3, o.field("cell3"))
    assert np.array_equal(c3[:, 0] + 16 * 16 * c3[:, 1] + 16 * c3[:, 2], cell1)


def test_rest_lattice_density_and_counts(small_dfsph):
    o = small_dfsph
    o.phase("compute_all_rho")
    o.phase("neighbour_counts")
    rho, cnt = o.field("rho"), o.field("nbr_count")
    # interior particle at rest spacing: 0.001 + 681.66 (self term excluded, SURVEY B-3); 26..32 neighbours
    # because the 6 two-spacing pairs sit at r ~ h within an ulp (App. A-7)
    assert abs(np.median(rho) - 681.66) < 2.0
    assert cnt.max() <= 32 and np.median(cnt) >= 26


def test_boundary_volume_brute_force(small_dfsph):
    # Akinci volume (PS:309-320) against an O(Nb^2) restatement in numpy fp32 for a few particles
    o = small_dfsph
    bpos, bvol = o.field("bpos"), o.field("bvol")
    h = np.float32(0.1)
    for i in (0, 17, 4000, 9001):
        d = bpos[i] - bpos
        r = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        m = (r <= h)
        m[i] = False
        s = sum(O.cubic_kernel(float(x), h) for x in r[m])
        assert abs(1.0 / s - bvol[i]) < 1e-5 * bvol[i]


def test_dfsph_first_step_is_free_fall(small_dfsph):
    # zero velocity => zero divergence error, no pressure; v = dt * g / m (SURVEY B-5: gravity / particle_m)
    o = O.Oracle(scenes.shipped("small_block", "dfsph"), threads=4)
    o.step()
    assert int(o.scalar("df_div_iters")) == 0 and int(o.scalar("df_den_iters")) == 2
    vel = o.field("vel")
    expect = -(1e-3 * 9.8 / 0.125) * 0.9999
    # particles with a symmetric neighbourhood feel no cohesion force: pure (reference-style) gravity
    interior = np.abs(vel[:, 1] - expect) < 1e-3 * abs(expect)
    assert interior.sum() > 1000
    assert abs(np.median(vel[:, 1]) - expect) < 5e-3 * abs(expect)
    o.close()


@pytest.mark.parametrize("solver", ["wcsph", "pcisph", "iisph", "dfsph"])
def test_solvers_run_and_stay_finite(solver):
    o = O.Oracle(scenes.shipped("small_block", solver), threads=4)
    for _ in range(3):
        o.step()
    assert np.isfinite(o.field("pos")).all() and np.isfinite(o.field("vel")).all()
    assert int(o.scalar("error_flags")) == 0
    # gravity acts: the block moves down
    assert o.field("vel")[:, 1].mean() < 0
    o.close()


def test_threads_do_not_change_results():
    a = O.Oracle(scenes.shipped("small_block", "dfsph"), threads=1)
    for _ in range(2):
        a.step()
    pa, va = a.field("pos").copy(), a.field("vel").copy()
    a.close()
    b = O.Oracle(scenes.shipped("small_block", "dfsph"), threads=4)
    for _ in range(2):
        b.step()
    assert np.array_equal(pa, b.field("pos")) and np.array_equal(va, b.field("vel"))
    b.close()


# ---------------------------------------------------------------------------------------------
# PBF (index-based semantics, oracle/sph_oracle_pbf.inc)
# ---------------------------------------------------------------------------------------------
def _poly6(r, h):
    q = r / h
    return np.where(q <= 1.0, 315.0 / (64.0 * np.pi * h ** 3) * (1.0 - q * q) ** 3, 0.0)


def test_pbf_density_lambda_against_brute_force():
    cfg = scenes.shipped("small_block", "pbf")
    o = O.Oracle(cfg, solver="pbf", threads=4)
    pos = o.field("pos").astype(np.float64)
    c = pos.mean(axis=0, keepdims=True)
    o.field("pos")[:] = (c + (pos - c) * 0.86).astype(np.float32)      # compress: constraint active
    o.base_step()
    o.phase("pbf_externel_force_predict_pos")
    o.phase("pbf_compute_all_lambda")
    pos = o.field("pos").astype(np.float64)
    bpos, bvol = o.field("bpos").astype(np.float64), o.field("bvol").astype(np.float64)
    h, m = 0.1, 1000 * 0.025 ** 3 * 8
    rho, lam, con = o.field("rho"), o.field("pbf_lambda"), o.field("pbf_constrain")
    for i in (0, 17, 2500, 5878):
        r = np.linalg.norm(pos - pos[i], axis=1)
        r[i] = np.inf                                                   # PS:461 self excluded
        rb = np.linalg.norm(bpos - pos[i], axis=1)
        brute = 0.001 + m * _poly6(r[r <= h + 1e-9], h).sum() + 1000.0 * (bvol * _poly6(rb, h))[rb <= h + 1e-9].sum()
        assert abs(brute - rho[i]) <= 2e-4 * brute, (i, brute, rho[i])
        assert con[i] == np.float32(max(np.float32(rho[i]) / np.float32(1000.0) - np.float32(1.0), 0.0))
        assert (lam[i] == 0.0) == (con[i] == 0.0) and lam[i] <= 0.0     # PBF:39-52
    assert (lam < 0).sum() > 100
    o.close()


def test_pbf_update_order_modes():
    # mode 1 = the reference's literal racy loop in ascending index order (ti.cpu, one thread); mode 0 =
    # move-all-then-XSPH, the order the CUDA path reproduces.  They differ only through the XSPH term.
    cfg = scenes.shipped("small_block", "pbf")
    out = []
    for mode in (0, 1):
        o = O.Oracle(cfg, solver="pbf", threads=4)
        o.set_scalar("pbf_update_mode", mode)
        rng = np.random.default_rng(5)
        o.field("vel")[:] = rng.normal(0, 0.5, o.field("vel").shape).astype(np.float32)
        o.step()
        out.append((o.field("pos").copy(), o.field("vel").copy()))
        o.close()
    assert np.array_equal(out[0][0], out[1][0])                         # positions of one step: identical
    dv = np.abs(out[0][1] - out[1][1]).max()
    assert 0.0 < dv < 0.2 * np.abs(out[0][1]).max()


def test_pbf_kernel_known_answers():
    h = 0.1
    # SB:122-129 poly6: W(0) = 315 / (64 pi h^3), W(h) = 0, unit integral over the support
    assert abs(O.poly_kernel(0.0, h) - 315.0 / (64.0 * math.pi * h ** 3)) <= 1e-3
    assert O.poly_kernel(h, h) == 0.0 and O.poly_kernel(1.5 * h, h) == 0.0
    r = (np.arange(4000) + 0.5) * (h / 4000)
    integral = sum(O.poly_kernel(x, h) * 4.0 * math.pi * x * x for x in r) * (h / 4000)
    assert abs(integral - 1.0) <= 2e-3
    # SB:113-120 spiky gradient: -45 (1 - q)^2 / (pi h^4) along r, zero at r = 0 and beyond the support
    g = O.spiky_kernel_derivative([0.05, 0.0, 0.0], h)
    assert abs(g[0] - (-45.0 * 0.25 / (math.pi * h ** 4))) <= 1e-3 * abs(g[0]) and g[1] == 0.0 and g[2] == 0.0
    assert not O.spiky_kernel_derivative([0.0, 0.0, 0.0], h).any()
    assert not O.spiky_kernel_derivative([0.0, 0.11, 0.0], h).any()
    d = O.spiky_kernel_derivative([0.03, -0.04, 0.0], h)              # |r| = 0.05: same magnitude, direction r / |r|
    assert abs(np.linalg.norm(d) - abs(g[0])) <= 1e-3 * abs(g[0]) and d[0] < 0 < d[1]


def _cubic_dw64(r, h):
    """SB:90-103 in float64, including the reference's extra factor 6."""
    rn = np.linalg.norm(r, axis=-1, keepdims=True)
    q = rn / h
    k = 48.0 / (np.pi * h ** 3)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = k * 6 * (3 * q * q - 2 * q) * r / (h * rn)
        b = -k * 6 * (1 - q) ** 2 * r / (h * rn)
    return np.where((q > 1e-5) & (q <= 0.5), a, np.where((q > 0.5) & (q <= 1.0), b, 0.0))


def test_dfsph_alpha_and_drho_against_float64_brute_force():
    # independent restatement of DF:32-89 (alpha) and DF:252-300 (D rho / D t) in numpy float64, all pairs,
    # on a jittered block with random velocities; the fp32 oracle must agree to fp32 accuracy
    cfg = scenes.shipped("small_block", "dfsph")
    o = O.Oracle(cfg, solver="dfsph", threads=4)
    rng = np.random.default_rng(2)
    o.field("pos")[:] += rng.uniform(-0.006, 0.006, o.field("pos").shape).astype(np.float32)
    o.field("vel")[:] = rng.normal(0, 1.0, o.field("vel").shape).astype(np.float32)
    o.base_step()
    o.phase("initialize")                      # rho, alpha
    o.phase("derivative_iter_all_rho")         # rho_derivative (+ neighbour counts)
    pos, vel = o.field("pos").astype(np.float64), o.field("vel").astype(np.float64)
    bpos, bvol = o.field("bpos").astype(np.float64), o.field("bvol").astype(np.float64)
    h, m = 0.1, 1000 * 0.025 ** 3 * 8
    h32 = float(np.float32(h))
    rho, alpha, drho, cnt = o.field("rho"), o.field("alpha"), o.field("rho_derivative"), o.field("nbr_count")
    checked = 0
    for i in rng.choice(len(pos), size=40, replace=False):
        r = pos[i] - pos
        d = np.linalg.norm(r, axis=1)
        nb = (d <= h32 * (1 + 1e-6)) & (np.arange(len(pos)) != i)
        if np.any(np.abs(d[nb] - h32) < 2e-6) or np.any(np.abs(np.linalg.norm(pos[i] - bpos, axis=1) - h32) < 2e-6):
            continue                            # a pair on the cut-off shell: fp32 and fp64 may disagree on membership
        g = m * _cubic_dw64(r[nb], h)
        rb = pos[i] - bpos
        nbb = np.linalg.norm(rb, axis=1) <= h32
        gb = (bvol[nbb, None] * 1000.0) * _cubic_dw64(rb[nbb], h)
        den = g.sum(0) @ g.sum(0) + (g * g).sum() + (gb * gb).sum() + gb.sum(0) @ gb.sum(0)      # DF:45
        want_alpha = 0.0 if abs(den) < 1e-6 else rho[i] / den
        assert abs(alpha[i] - want_alpha) <= 2e-4 * abs(want_alpha) + 1e-12, (i, alpha[i], want_alpha)
        assert cnt[i] == nb.sum()
        if cnt[i] >= 20:                                                                          # DF:258-261
            dr = m * ((vel[i] - vel[nb]) * _cubic_dw64(r[nb], h)).sum() + 1000.0 * (
                bvol[nbb] * (_cubic_dw64(rb[nbb], h) @ vel[i])).sum()                             # DF:287, 300, 267
            want = max(dr, 0.0)
            scale = m * np.abs((vel[i] - vel[nb]) * _cubic_dw64(r[nb], h)).sum() + 1.0
            assert abs(drho[i] - want) <= 2e-5 * scale, (i, drho[i], want)
        else:
            assert drho[i] == 0.0
        checked += 1
    assert checked >= 25
    o.close()


def _cubic_w64(r, h):
    """SB:74-88 in float64."""
    q = r / h
    k = 8.0 / (np.pi * h ** 3)
    return np.where(q <= 0.5, k * (6 * (q ** 3 - q ** 2) + 1), np.where(q <= 1.0, 2 * k * (1 - q) ** 3, 0.0))


def test_wcsph_pressure_phase_against_float64_brute_force():
    # independent restatement of WC:65-129 (Tait pressure, symmetric pressure gradient, Akinci boundary term) and of
    # SB:41-72 (density) in numpy float64 over all pairs, on a compressed block so that pressures are non-zero
    cfg = scenes.shipped("small_block", "wcsph")
    o = O.Oracle(cfg, solver="wcsph", threads=4)
    rng = np.random.default_rng(4)
    pos = o.field("pos").astype(np.float64)
    c = pos.mean(axis=0, keepdims=True)
    o.field("pos")[:] = (c + (pos - c) * 0.85 + rng.uniform(-0.003, 0.003, pos.shape)).astype(np.float32)   # rest rho is ~690
    o.base_step()
    o.phase("pressure_phase")
    pos = o.field("pos").astype(np.float64)
    bpos, bvol = o.field("bpos").astype(np.float64), o.field("bvol").astype(np.float64)
    h, m = 0.1, 1000 * 0.025 ** 3 * 8
    h32 = float(np.float32(h))
    rho, p = o.field("rho").astype(np.float64), o.field("pressure").astype(np.float64)
    pg, ba = o.field("pressure_gradient"), o.field("boundary_acc")
    assert (p > 0).sum() > 1000
    # Tait equation for every particle (WC:86-90)
    want_p = 70000.0 * ((np.maximum(rho, 1000.0) / 1000.0) ** 7 - 1.0)
    assert np.allclose(p, want_p, rtol=2e-5, atol=0.05)
    checked = 0
    for i in rng.choice(len(pos), size=40, replace=False):
        r = pos[i] - pos
        d = np.linalg.norm(r, axis=1)
        nb = (d <= h32) & (np.arange(len(pos)) != i)
        db = np.linalg.norm(pos[i] - bpos, axis=1)
        if np.any(np.abs(d[nb] - h32) < 2e-6) or np.any(np.abs(db - h32) < 2e-6):
            continue
        nbb = db <= h32
        want_rho = 0.001 + m * _cubic_w64(d[nb], h).sum() + 1000.0 * (bvol[nbb] * _cubic_w64(db[nbb], h)).sum()   # SB:44-49
        assert abs(rho[i] - want_rho) <= 1e-5 * want_rho
        dw = _cubic_dw64(r[nb], h)
        want_pg = -(m * (p[i] / rho[i] ** 2 + p[nb] / rho[nb] ** 2)[:, None] * dw).sum(0)                          # WC:116
        scale = np.abs(m * (p[i] / rho[i] ** 2 + p[nb] / rho[nb] ** 2)[:, None] * dw).sum() + 1e-3
        assert np.abs(pg[i] - want_pg).max() <= 2e-5 * scale, (i, pg[i], want_pg)
        want_ba = -1000.0 * ((bvol[nbb] * p[i] / rho[i] ** 2)[:, None] * _cubic_dw64((pos[i] - bpos)[nbb], h)).sum(0)  # WC:83, 99
        assert np.abs(ba[i] - want_ba).max() <= 2e-5 * (np.abs(want_ba).max() + 1e-3)
        checked += 1
    assert checked >= 25
    o.close()


def test_iisph_predict_advection_against_float64_brute_force():
    # independent restatement of II:35-75 + the tasks II:255-340 (d_ii, rho_adv, a_ii) in numpy float64, all pairs
    cfg = scenes.shipped("small_block", "iisph")
    o = O.Oracle(cfg, solver="iisph", threads=4)
    rng = np.random.default_rng(6)
    o.field("pos")[:] += rng.uniform(-0.006, 0.006, o.field("pos").shape).astype(np.float32)
    o.field("vel")[:] = rng.normal(0, 0.5, o.field("vel").shape).astype(np.float32)
    o.base_step()
    o.phase("ii_predict_advection")
    pos = o.field("pos").astype(np.float64)
    bpos, bvol = o.field("bpos").astype(np.float64), o.field("bvol").astype(np.float64)
    h, m, dt = 0.1, 1000 * 0.025 ** 3 * 8, float(np.float32(o.scalar("delta_time")))
    h32 = float(np.float32(h))
    rho, v_adv = o.field("rho").astype(np.float64), o.field("v_adv").astype(np.float64)
    d_ii, a_ii, rho_adv = o.field("d_ii"), o.field("a_ii"), o.field("rho_adv")
    checked = 0
    for i in rng.choice(len(pos), size=40, replace=False):
        r = pos[i] - pos
        d = np.linalg.norm(r, axis=1)
        nb = (d <= h32) & (np.arange(len(pos)) != i)
        rb = pos[i] - bpos
        db = np.linalg.norm(rb, axis=1)
        if np.any(np.abs(d[nb] - h32) < 2e-6) or np.any(np.abs(db - h32) < 2e-6):
            continue
        nbb = db <= h32
        dw, dwb = _cubic_dw64(r[nb], h), _cubic_dw64(rb[nbb], h)
        want_dii = ((-m / rho[i] ** 2) * dw.sum(0) + 1000.0 * ((-bvol[nbb] / rho[i] ** 2)[:, None] * dwb).sum(0)) * dt * dt  # II:53
        sc = (np.abs((m / rho[i] ** 2) * dw).sum() + 1000.0 * np.abs((bvol[nbb] / rho[i] ** 2)[:, None] * dwb).sum()) * dt * dt
        assert np.abs(d_ii[i] - want_dii).max() <= 2e-5 * sc + 1e-20, (i, d_ii[i], want_dii)
        want_ra = (m * ((v_adv[i] - v_adv[nb]) * dw).sum() + 1000.0 * (bvol[nbb] * (dwb @ v_adv[i])).sum()) * dt + rho[i]     # II:63
        assert abs(rho_adv[i] - want_ra) <= 2e-5 * abs(want_ra)
        dji = (-dt * dt * m / rho[i] ** 2) * _cubic_dw64(-r[nb], h)                                                       # II:283-284
        djib = (-dt * dt * m / rho[i] ** 2) * _cubic_dw64(-rb[nbb], h)
        dii = d_ii[i].astype(np.float64)
        want_aii = m * ((dii - dji) * dw).sum() + 1000.0 * (bvol[nbb] * ((dii - djib) * dwb).sum(1)).sum()                # II:73, 285, 303
        sca = m * np.abs((dii - dji) * dw).sum() + 1000.0 * np.abs(bvol[nbb][:, None] * (dii - djib) * dwb).sum()
        assert abs(a_ii[i] - want_aii) <= 3e-5 * sca + 1e-20, (i, a_ii[i], want_aii)
        checked += 1
    assert checked >= 25
    o.close()


def test_pcisph_delta_predicted_density_and_force_against_float64_brute_force():
    # independent restatement of PC:39-45 (delta), PC:89-101 (predicted density, neighbour set from the CURRENT
    # positions, kernel on the PREDICTED ones) and PC:109-119 (pressure force) in numpy float64, all pairs
    cfg = scenes.shipped("small_block", "pcisph")
    o = O.Oracle(cfg, solver="pcisph", threads=4)
    pos0 = o.field("pos").astype(np.float64)
    h, m = 0.1, 1000 * 0.025 ** 3 * 8
    h32 = float(np.float32(h))
    dt = float(np.float32(cfg["solver"]["delta_time"]))
    k = int(o.scalar("pc_max_index"))
    r = pos0[k] - pos0
    d = np.linalg.norm(r, axis=1)
    nb = (d <= h32 * (1 + 2e-6)) & (np.arange(len(pos0)) != k)           # the lattice has pairs exactly on the shell (W' = 0 there)
    g = _cubic_dw64(r[nb], h)
    beta = dt * dt * m * m * 2 / 1000.0 ** 2                             # PC:23
    want_delta = 1.0 / ((g.sum(0) @ g.sum(0) + (g * g).sum()) * beta)     # PC:45
    assert abs(o.scalar("pc_delta") - want_delta) <= 2e-4 * want_delta
    # a compressed, jittered state so that the pressure loop runs
    rng = np.random.default_rng(8)
    c = pos0.mean(axis=0, keepdims=True)
    o.field("pos")[:] = (c + (pos0 - c) * 0.86 + rng.uniform(-0.003, 0.003, pos0.shape)).astype(np.float32)
    o.base_step()
    o.phase("pc_compute_ext_force")
    o.phase("pc_iteration")
    assert int(o.scalar("pc_iters")) >= 1
    pos, pp = o.field("pos").astype(np.float64), o.field("pos_predict").astype(np.float64)
    bpos, bvol = o.field("bpos").astype(np.float64), o.field("bvol").astype(np.float64)
    rho, press = o.field("rho").astype(np.float64), o.field("press_iter").astype(np.float64)
    rho_predict, pf = o.field("rho_predict"), o.field("press_force")
    checked = 0
    for i in rng.choice(len(pos), size=40, replace=False):
        r = pos[i] - pos
        d = np.linalg.norm(r, axis=1)
        nb = (d <= h32) & (np.arange(len(pos)) != i)
        db = np.linalg.norm(pos[i] - bpos, axis=1)
        if np.any(np.abs(d[nb] - h32) < 2e-6) or np.any(np.abs(db - h32) < 2e-6):
            continue
        nbb = db <= h32
        want_rp = m * _cubic_w64(np.linalg.norm(pp[i] - pp[nb], axis=1), h).sum() + 1000.0 * (
            bvol[nbb] * _cubic_w64(np.linalg.norm(pp[i] - bpos[nbb], axis=1), h)).sum()            # PC:98, 141-153
        assert abs(rho_predict[i] - want_rp) <= 2e-5 * want_rp
        # press_force belongs to the last executed force pass; press_iter still holds the pressures it used
        dw = _cubic_dw64(r[nb], h)
        f = ((press[i] + press[nb])[:, None] * dw / 1000.0 ** 2 * m * m).sum(0)                    # PC:177
        fb = -((bvol[nbb] * press[i] / rho[i] ** 2)[:, None] * _cubic_dw64((pos[i] - bpos)[nbb], h)).sum(0)   # PC:197
        want_pf = -f + fb * 1000.0 * m                                                             # PC:117
        sc = np.abs((press[i] + press[nb])[:, None] * dw / 1000.0 ** 2 * m * m).sum() + 1e-9
        assert np.abs(pf[i] - want_pf).max() <= 3e-5 * sc + 1e-12, (i, pf[i], want_pf)
        checked += 1
    assert checked >= 25
    o.close()


def test_viscosity_tension_and_divergence_update_against_float64_brute_force():
    # independent restatement of SB:170-217 (artificial viscosity, cohesion) and of one divergence iteration
    # DF:302-312, 357-391 (velocity correction from kappa = D rho/D t * alpha / dt) in numpy float64, all pairs
    cfg = scenes.shipped("small_block", "dfsph")
    o = O.Oracle(cfg, solver="dfsph", threads=4)
    rng = np.random.default_rng(12)
    o.field("pos")[:] += rng.uniform(-0.006, 0.006, o.field("pos").shape).astype(np.float32)
    o.field("vel")[:] = rng.normal(0, 1.0, o.field("vel").shape).astype(np.float32)
    o.base_step()
    o.phase("initialize")
    o.phase("solve_all_viscosity")
    o.phase("solve_all_tension")
    o.phase("derivative_iter_all_rho")
    vel0 = o.field("vel").astype(np.float64).copy()
    o.phase("divergence_iter_all_vel_adv")
    pos, vel1 = o.field("pos").astype(np.float64), o.field("vel").astype(np.float64)
    bpos, bvol = o.field("bpos").astype(np.float64), o.field("bvol").astype(np.float64)
    rho, alpha, drho = (o.field(k).astype(np.float64) for k in ("rho", "alpha", "rho_derivative"))
    visc, ten = o.field("viscosity"), o.field("tension")
    h, m, dt = 0.1, 1000 * 0.025 ** 3 * 8, float(np.float32(o.scalar("delta_time")))
    h32 = float(np.float32(h))
    kappa = drho * alpha / dt
    checked = 0
    for i in rng.choice(len(pos), size=40, replace=False):
        r = pos[i] - pos
        d = np.linalg.norm(r, axis=1)
        nb = (d <= h32) & (np.arange(len(pos)) != i)
        rb = pos[i] - bpos
        db = np.linalg.norm(rb, axis=1)
        if np.any(np.abs(d[nb] - h32) < 2e-6) or np.any(np.abs(db - h32) < 2e-6):
            continue
        nbb = db <= h32
        dw = _cubic_dw64(r[nb], h)
        # SB:204-217 tension = m * sum(-k_t / m * m * W * x_ij), k_t = 0.5
        want_t = m * ((-0.5 / m * m) * _cubic_w64(d[nb], h)[:, None] * r[nb]).sum(0)
        assert np.abs(ten[i] - want_t).max() <= 2e-5 * (np.abs(m * 0.5 * _cubic_w64(d[nb], h)[:, None] * r[nb]).sum() + 1e-9)
        # SB:170-189 viscosity, alpha = 0.08, c_s = 13, eps = 0.01, only approaching pairs
        shear = ((vel0[i] - vel0[nb]) * r[nb]).sum(1)
        nu = (2 * 0.08 * h * 13) / (rho[i] + rho[nb])
        pi_ij = -nu * shear / (d[nb] ** 2 + 0.01 * h * h)
        terms = np.where((shear < 0)[:, None], -m * pi_ij[:, None] * dw, 0.0)
        want_v = m * terms.sum(0)
        assert np.abs(visc[i] - want_v).max() <= 3e-5 * (np.abs(m * terms).sum() + 1e-9), (i, visc[i], want_v)
        # DF:302-312 one divergence iteration: v -= dt * (sum_j m (k_i/rho_i + k_j/rho_j) grad W [if > 1e-5] + rho0 * sum_b ...)
        s = kappa[i] / rho[i] + kappa[nb] / rho[nb]
        if np.any(np.abs(s - 1e-5) < 1e-6):
            continue                              # a pair on the threshold of DF:367
        acc = (m * np.where(s > 1e-5, s, 0.0)[:, None] * dw).sum(0) + 1000.0 * (
            (bvol[nbb] * kappa[i] / rho[i])[:, None] * _cubic_dw64(rb[nbb], h)).sum(0)
        want_vel = vel0[i] - acc * dt
        sc = dt * (np.abs(m * s[:, None] * dw).sum() + 1e-9) + np.abs(vel0[i]).max() * 1e-2
        assert np.abs(vel1[i] - want_vel).max() <= 3e-5 * sc, (i, vel1[i], want_vel)
        checked += 1
    assert checked >= 20
    o.close()


def test_rigid_mass_properties_against_float64_brute_force():
    # PS:249-307: V_r = 1 / sum_{r' != r} W, m_r = rho_r V_r, centroid, inertia tensor and its inverse, in float64
    lo, hi, pitch = [0, 0, 0], [0.3, 0.4, 0.5], 0.05
    ax = [np.arange(int(round(lo[k] / pitch)), int(round(hi[k] / pitch)) + 1) * pitch for k in range(3)]
    pts = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    verts = np.array([[x, y, z] for x in (0.0, 0.3) for y in (0.0, 0.4) for z in (0.0, 0.5)], dtype=np.float32)
    cfg = scenes.make_scene([2.0, 2.0, 1.0], [0.1, 0.1, 0.1], [0.6, 0.8, 0.8], "dfsph", 1e-4,
                            solid={"mesh": "unused", "voxel_radius": 0.025, "rho_0": 2000, "scale": 1,
                                   "pos_offset": [0.85, 0.0, 0.2], "attitude_offset": [0.0, 0.0, 0.0],
                                   "fill": True, "active": True})
    o = O.Oracle(cfg, solver="dfsph", rigid_points=pts, rigid_vertices=verts, threads=1)
    rp = o.field("rpos").astype(np.float64)
    assert np.allclose(rp, pts.astype(np.float64) + np.array([0.85, 0.0, 0.2]), atol=1e-6)       # PS:198-223, no rotation
    d = np.linalg.norm(rp[:, None, :] - rp[None, :, :], axis=2)
    w = _cubic_w64(d, 0.1)
    w[(d > float(np.float32(0.1))) | np.eye(len(rp), dtype=bool)] = 0.0
    vol = 1.0 / w.sum(1)
    assert np.allclose(o.field("rvol"), vol, rtol=1e-5)
    mass = 2000.0 * vol
    assert np.allclose(o.field("rmass"), mass, rtol=1e-5)
    cen = (rp * mass[:, None]).sum(0) / mass.sum()
    assert np.allclose(o.field("centroid").reshape(-1), cen, atol=2e-6)
    q = rp - cen
    I = np.zeros((3, 3))
    I[0, 0] = (mass * (q[:, 1] ** 2 + q[:, 2] ** 2)).sum()
    I[1, 1] = (mass * (q[:, 0] ** 2 + q[:, 2] ** 2)).sum()
    I[2, 2] = (mass * (q[:, 0] ** 2 + q[:, 1] ** 2)).sum()
    I[0, 1] = I[1, 0] = -(mass * q[:, 0] * q[:, 1]).sum()
    I[0, 2] = I[2, 0] = -(mass * q[:, 0] * q[:, 2]).sum()
    I[1, 2] = I[2, 1] = -(mass * q[:, 2] * q[:, 1]).sum()
    got = o.field("inertia").reshape(3, 3).astype(np.float64)
    assert np.allclose(np.diag(got), np.diag(I), rtol=2e-5)
    assert np.abs(got - I).max() <= 2e-5 * np.abs(np.diag(I)).max()
    assert np.allclose(o.field("inertia_inv").reshape(3, 3) @ got, np.eye(3), atol=1e-4)
    o.close()


def test_dfsph_density_iteration_against_float64_brute_force():
    # independent restatement of DF:124-152 (rho_adv = max(rho + dt * div, rho0)) and DF:178-219 (one constant-density
    # iteration of v*) in numpy float64, all pairs, on a compressed block so that rho_adv > rho0 for many particles
    cfg = scenes.shipped("small_block", "dfsph")
    o = O.Oracle(cfg, solver="dfsph", threads=4)
    rng = np.random.default_rng(14)
    pos0 = o.field("pos").astype(np.float64)
    c = pos0.mean(axis=0, keepdims=True)
    o.field("pos")[:] = (c + (pos0 - c) * 0.86 + rng.uniform(-0.003, 0.003, pos0.shape)).astype(np.float32)
    o.field("vel")[:] = rng.normal(0, 0.5, pos0.shape).astype(np.float32)
    o.base_step()
    o.phase("initialize")
    o.phase("compute_all_ext_force")
    o.phase("compute_all_vel_adv")
    o.phase("compute_all_rho_adv")
    va0 = o.field("vel_adv").astype(np.float64).copy()
    o.phase("iter_all_vel_adv")
    pos, va1 = o.field("pos").astype(np.float64), o.field("vel_adv").astype(np.float64)
    bpos, bvol = o.field("bpos").astype(np.float64), o.field("bvol").astype(np.float64)
    rho, alpha, rho_adv = (o.field(k).astype(np.float64) for k in ("rho", "alpha", "rho_adv"))
    h, m = 0.1, 1000 * 0.025 ** 3 * 8
    h32 = float(np.float32(h))
    dt, dt2 = float(np.float32(o.scalar("delta_time"))), float(np.float32(o.scalar("delta_time_2")))
    assert (rho_adv > 1000.0).sum() > 500
    k = (rho_adv - 1000.0) * alpha / dt2
    checked = 0
    for i in rng.choice(len(pos), size=40, replace=False):
        r = pos[i] - pos
        d = np.linalg.norm(r, axis=1)
        nb = (d <= h32) & (np.arange(len(pos)) != i)
        rb = pos[i] - bpos
        db = np.linalg.norm(rb, axis=1)
        if np.any(np.abs(d[nb] - h32) < 2e-6) or np.any(np.abs(db - h32) < 2e-6):
            continue
        nbb = db <= h32
        dw, dwb = _cubic_dw64(r[nb], h), _cubic_dw64(rb[nbb], h)
        div = m * ((va0[i] - va0[nb]) * dw).sum() + 1000.0 * (bvol[nbb] * (dwb @ va0[i])).sum()
        want_ra = max(rho[i] + dt * div, 1000.0)                                                    # DF:135
        sc = dt * (m * np.abs((va0[i] - va0[nb]) * dw).sum() + 1.0)
        assert abs(rho_adv[i] - want_ra) <= 2e-5 * (abs(want_ra) + sc), (i, rho_adv[i], want_ra)
        delta = (m * (k[i] / rho[i] + k[nb] / rho[nb])[:, None] * dw).sum(0) + 1000.0 * (
            (bvol[nbb] * k[i] / rho[i])[:, None] * dwb).sum(0)                                       # DF:187, 203, 219
        want_va = va0[i] - delta * dt                                                               # DF:191
        sc = dt * np.abs(m * (k[i] / rho[i] + k[nb] / rho[nb])[:, None] * dw).sum() + 1e-2 * np.abs(va0[i]).max() + 1e-9
        assert np.abs(va1[i] - want_va).max() <= 3e-5 * sc, (i, va1[i], want_va)
        checked += 1
    assert checked >= 25
    o.close()


def test_iisph_relaxed_jacobi_update_against_float64_brute_force():
    # independent restatement of II:121-147 + II:228-253 (sum_j d_ij p_j, the r_sum of the relaxed Jacobi update,
    # omega = 0.5, clamp at 0) in numpy float64, all pairs, from a non-trivial pressure iterate
    cfg = scenes.shipped("small_block", "iisph")
    o = O.Oracle(cfg, solver="iisph", threads=4)
    rng = np.random.default_rng(16)
    pos0 = o.field("pos").astype(np.float64)
    c = pos0.mean(axis=0, keepdims=True)
    o.field("pos")[:] = (c + (pos0 - c) * 0.86 + rng.uniform(-0.003, 0.003, pos0.shape)).astype(np.float32)
    o.field("p_past")[:] = rng.uniform(0.0, 4000.0, len(pos0)).astype(np.float32)
    o.base_step()
    o.phase("ii_predict_advection")            # p_iter = 0.5 p_past, d_ii, a_ii, rho_adv
    p0 = o.field("p_iter").astype(np.float64).copy()
    assert np.allclose(p0, 0.5 * o.field("p_past"))
    o.phase("ii_compute_all_d_ij")
    o.phase("ii_update_p")
    pos = o.field("pos").astype(np.float64)
    bpos, bvol = o.field("bpos").astype(np.float64), o.field("bvol").astype(np.float64)
    rho, a_ii, rho_adv = (o.field(k).astype(np.float64) for k in ("rho", "a_ii", "rho_adv"))
    d_ii, d_ij = o.field("d_ii").astype(np.float64), o.field("d_ij").astype(np.float64)
    r_sum, p1 = o.field("r_sum"), o.field("p_iter")
    h, m, dt = 0.1, 1000 * 0.025 ** 3 * 8, float(np.float32(o.scalar("delta_time")))
    h32 = float(np.float32(h))
    checked = 0
    for i in rng.choice(len(pos), size=40, replace=False):
        r = pos[i] - pos
        d = np.linalg.norm(r, axis=1)
        nb = (d <= h32) & (np.arange(len(pos)) != i)
        rb = pos[i] - bpos
        db = np.linalg.norm(rb, axis=1)
        if np.any(np.abs(d[nb] - h32) < 2e-6) or np.any(np.abs(db - h32) < 2e-6):
            continue
        nbb = db <= h32
        dw, dwb = _cubic_dw64(r[nb], h), _cubic_dw64(rb[nbb], h)
        want_dij = (-m * (p0[nb] / rho[nb] ** 2)[:, None] * dw).sum(0) * dt * dt                     # II:126, 313
        sc = np.abs(m * (p0[nb] / rho[nb] ** 2)[:, None] * dw).sum() * dt * dt + 1e-30
        assert np.abs(d_ij[i] - want_dij).max() <= 2e-5 * sc, (i, d_ij[i], want_dij)
        d_ji = (-dt * dt * m / rho[i] ** 2) * _cubic_dw64(-r[nb], h) * p0[i]                         # II:244-245
        t = d_ij[i] - d_ii[nb] * p0[nb][:, None] - (d_ij[nb] - d_ji)                                 # II:246
        want_rs = m * (t * dw).sum() + 1000.0 * (bvol[nbb] * (dwb @ d_ij[i])).sum()                  # II:136, 232
        scr = m * np.abs(t * dw).sum() + 1000.0 * np.abs(bvol[nbb][:, None] * dwb * d_ij[i]).sum() + 1e-30
        assert abs(r_sum[i] - want_rs) <= 5e-5 * scr, (i, r_sum[i], want_rs)
        want_p = 0.0
        if abs(a_ii[i]) > 1e-7:
            want_p = 0.5 * p0[i] + 0.5 * (1000.0 - rho_adv[i] - float(r_sum[i])) / a_ii[i]           # II:140-142
        want_p = max(want_p, 0.0)                                                                    # II:147
        assert abs(p1[i] - want_p) <= 2e-5 * (abs(want_p) + abs(0.5 * (1000.0 - rho_adv[i]) / a_ii[i]) + 1.0)
        checked += 1
    assert checked >= 25
    o.close()


def test_rigid_free_flight_known_answer():
    # RS:33-104, 216-234 with no fluid contact and no wall contact: acc = g (0, -1, 0), v += acc dt, every particle,
    # vertex and the centroid move by v dt; no torque, so omega stays zero and the body does not rotate
    lo, hi, pitch = [0, 0, 0], [0.2, 0.2, 0.2], 0.05
    ax = [np.arange(int(round(lo[k] / pitch)), int(round(hi[k] / pitch)) + 1) * pitch for k in range(3)]
    pts = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    verts = np.array([[x, y, z] for x in (0.0, 0.2) for y in (0.0, 0.2) for z in (0.0, 0.2)], dtype=np.float32)
    dt, g = 1e-3, 9.8
    cfg = scenes.make_scene([3.0, 3.0, 1.5], [0.1, 0.1, 0.1], [0.3, 0.3, 0.3], "wcsph", dt,
                            solid={"mesh": "unused", "voxel_radius": 0.025, "rho_0": 2000, "scale": 1,
                                   "pos_offset": [2.0, 2.0, 0.6], "attitude_offset": [0.0, 0.0, 0.0],
                                   "fill": True, "active": True})
    o = O.Oracle(cfg, solver="wcsph", rigid_points=pts, rigid_vertices=verts, threads=1)
    p0, c0 = o.field("rpos").astype(np.float64).copy(), o.field("centroid").astype(np.float64).copy().reshape(-1)
    v, y = 0.0, 0.0
    for n in range(1, 6):
        o.step(1)                                  # fluid step (far away) + rigid step
        v -= g * dt
        y += v * dt
        assert np.allclose(o.field("rvel")[:, 1], v, rtol=1e-5) and not o.field("rvel")[:, [0, 2]].any()
        assert np.allclose(o.field("rpos") - p0, [0.0, y, 0.0], atol=2e-6)
        assert np.allclose(o.field("centroid").reshape(-1) - c0, [0.0, y, 0.0], atol=2e-6)
        assert not o.field("rs_omega").any()
    assert abs(o.scalar("rs_mass") - o.field("rmass").astype(np.float64).sum()) <= 1e-4 * o.scalar("rs_mass")   # RS:156-162
    o.close()


def test_fluid_to_rigid_force_against_float64_brute_force():
    # DF:204-212: every fluid-rigid pair adds  m * V_r rho0 k_i / rho_i * grad W(x_i - x_r)  to the rigid particle's
    # force during iter_all_vel_adv; restated per rigid particle in numpy float64 over all fluid particles
    lo, hi, pitch = [0, 0, 0], [0.2, 0.3, 0.2], 0.05
    ax = [np.arange(int(round(lo[k] / pitch)), int(round(hi[k] / pitch)) + 1) * pitch for k in range(3)]
    pts = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    verts = np.array([[x, y, z] for x in (0.0, 0.2) for y in (0.0, 0.3) for z in (0.0, 0.2)], dtype=np.float32)
    cfg = scenes.make_scene([2.0, 2.0, 1.0], [0.1, 0.1, 0.1], [0.6, 0.8, 0.8], "dfsph", 1e-4,
                            solid={"mesh": "unused", "voxel_radius": 0.025, "rho_0": 2000, "scale": 1,
                                   "pos_offset": [0.42, 0.3, 0.3], "attitude_offset": [0.0, 0.0, 0.0],   # inside the block: k_i > 0 around it
                                   "fill": True, "active": True})
    o = O.Oracle(cfg, solver="dfsph", rigid_points=pts, rigid_vertices=verts, threads=1)
    rng = np.random.default_rng(18)
    pos0 = o.field("pos").astype(np.float64)
    c = pos0.mean(axis=0, keepdims=True)
    o.field("pos")[:] = (c + (pos0 - c) * 0.86 + rng.uniform(-0.002, 0.002, pos0.shape)).astype(np.float32)
    o.base_step()
    for ph in ("initialize", "compute_all_ext_force", "compute_all_vel_adv", "compute_all_rho_adv"):
        o.phase(ph)
    assert not o.field("rforce").any()
    o.phase("iter_all_vel_adv")
    pos, rp, rvol = o.field("pos").astype(np.float64), o.field("rpos").astype(np.float64), o.field("rvol").astype(np.float64)
    rho, alpha, rho_adv = (o.field(k).astype(np.float64) for k in ("rho", "alpha", "rho_adv"))
    m, h = 1000 * 0.025 ** 3 * 8, 0.1
    h32 = float(np.float32(h))
    dt2 = float(np.float32(o.scalar("delta_time_2")))
    k = (rho_adv - 1000.0) * alpha / dt2                                                        # DF:208
    got = o.field("rforce").astype(np.float64)
    assert np.abs(got).max() > 0
    checked = 0
    for r in range(len(rp)):
        x = pos - rp[r]
        d = np.linalg.norm(x, axis=1)
        nb = d <= h32
        if not nb.any():
            assert not got[r].any()
            continue
        if np.any(np.abs(d[nb] - h32) < 2e-6):
            continue
        terms = m * (rvol[r] * 1000.0 * k[nb] / rho[nb])[:, None] * _cubic_dw64(x[nb], h)       # DF:211-212
        assert np.abs(got[r] - terms.sum(0)).max() <= 3e-5 * (np.abs(terms).sum() + 1e-12), (r, got[r], terms.sum(0))
        checked += 1
    assert checked >= 20
    o.close()


def _spiky_dw64(r, h):
    rn = np.linalg.norm(r, axis=-1, keepdims=True)
    q = rn / h
    with np.errstate(divide="ignore", invalid="ignore"):
        g = -(45.0 * (1 - q) ** 2) * r / (np.pi * h ** 4 * rn)
    return np.where((q <= 1.0) & (q > 0.0), g, 0.0)


def test_pbf_delta_pos_against_float64_brute_force():
    # PBF:55-65 + 164-174: delta_p_i = (sum_j (lambda_i + lambda_j + s_corr) grad W_spiky + sum_b (lambda_i + s_corr) grad W) / rho0,
    # s_corr = -k (W_poly6(r) / W_poly6(0.3 h))^4, restated in numpy float64 over all pairs
    cfg = scenes.shipped("small_block", "pbf")
    o = O.Oracle(cfg, solver="pbf", threads=4)
    rng = np.random.default_rng(20)
    pos0 = o.field("pos").astype(np.float64)
    c = pos0.mean(axis=0, keepdims=True)
    o.field("pos")[:] = (c + (pos0 - c) * 0.86 + rng.uniform(-0.002, 0.002, pos0.shape)).astype(np.float32)
    o.base_step()
    for ph in ("pbf_externel_force_predict_pos", "pbf_compute_all_lambda", "pbf_compute_all_delta_pos"):
        o.phase(ph)
    pos, lam = o.field("pos").astype(np.float64), o.field("pbf_lambda").astype(np.float64)
    bpos = o.field("bpos").astype(np.float64)
    dp = o.field("pbf_delta_pos")
    h = 0.1
    h32 = float(np.float32(h))
    w03 = float(_poly6(np.float64(np.float32(0.3 * 0.1)), h))
    checked = 0
    for i in rng.choice(len(pos), size=40, replace=False):
        r = pos[i] - pos
        d = np.linalg.norm(r, axis=1)
        nb = (d <= h32) & (np.arange(len(pos)) != i)
        rb = pos[i] - bpos
        db = np.linalg.norm(rb, axis=1)
        if np.any(np.abs(d[nb] - h32) < 2e-6) or np.any(np.abs(db - h32) < 2e-6):
            continue
        nbb = db <= h32
        sc = -1e-7 * (_poly6(d[nb], h) / w03) ** 4
        scb = -1e-7 * (_poly6(db[nbb], h) / w03) ** 4
        tf = (lam[i] + lam[nb] + sc)[:, None] * _spiky_dw64(r[nb], h)
        tb = (lam[i] + scb)[:, None] * _spiky_dw64(rb[nbb], h)
        want = (tf.sum(0) + tb.sum(0)) / 1000.0
        scale = (np.abs(tf).sum() + np.abs(tb).sum()) / 1000.0 + 1e-30
        assert np.abs(dp[i] - want).max() <= 3e-5 * scale, (i, dp[i], want)
        checked += 1
    assert checked >= 25
    o.close()


def test_dfsph_adaptive_time_step_known_answer():
    # DF:98-122: v* = v + dt f / m, dt_new = clamp(0.4 * d / max|v*| * 0.2, 1e-5, 1e-3), delta_time_2 = dt_new^2
    cfg = scenes.shipped("small_block", "dfsph")
    o = O.Oracle(cfg, solver="dfsph", threads=4)
    o.field("vel")[:, 0] = 3.0                   # 0.4 * 0.05 / 3 * 0.2 = 1.33e-3 -> clamped to 1e-3
    o.base_step(); o.phase("initialize"); o.phase("compute_all_ext_force"); o.phase("compute_all_vel_adv")
    assert np.float32(o.scalar("delta_time")) == np.float32(1e-3)
    o.field("vel")[:, 0] = 8.0
    o.base_step(); o.phase("initialize"); o.phase("compute_all_ext_force"); o.phase("compute_all_vel_adv")
    vmax = np.linalg.norm(o.field("vel_adv").astype(np.float64), axis=1).max()
    want = 0.4 * 0.05 / vmax * 0.2
    assert abs(o.scalar("delta_time") - want) <= 1e-5 * want and 1e-5 < want < 1e-3
    assert abs(o.scalar("delta_time_2") - want * want) <= 1e-5 * want * want
    assert abs(o.scalar("ps_delta_time") - want) <= 1e-5 * want             # DF:119, read by the rigid solver
    o.close()
