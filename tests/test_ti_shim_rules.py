"""The rules of tests/golden/ti_shim/taichi (the stand-in that executes the reference's sources for the refshim_* fixtures),
one small kernel per rule, so that what the fixtures rest on is stated and checked: SURVEY.md appendix A, item by item."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


ti = None      # the stand-in, while a test of this module runs (kernels look `ti` up in the module's globals)


@pytest.fixture(scope="module", autouse=True)
def shim():
    global ti
    saved = sys.modules.pop("taichi", None)
    sys.path.insert(0, os.path.join(HERE, "golden", "ti_shim"))
    try:
        import taichi
        assert taichi.__file__.startswith(os.path.join(HERE, "golden", "ti_shim"))
        ti = taichi
        yield taichi
    finally:
        ti = None
        sys.path.remove(os.path.join(HERE, "golden", "ti_shim"))
        sys.modules.pop("taichi", None)
        if saved is not None:
            sys.modules["taichi"] = saved


F = np.float32


def test_constants_fold_in_binary64_and_meet_runtime_values_as_binary32():
    m = 1000 * (0.025 ** 3) * 8                      # ParticleSystem.py:83 -> 0.12500000000000003

    class K:
        eps, h, m = 0.01, 0.1, 1000 * (0.025 ** 3) * 8
        out = ti.field(ti.f32, shape=4)

        @ti.kernel
        def run(self):
            a = 1.0 / 3.0                            # constant / constant: Python, binary64, then an f32 local
            self.out[0] = a
            q = self.out[0] * 3                      # runtime f32 (x) int constant
            self.out[1] = q + self.eps * self.h * self.h      # constants fold left to right in binary64 first (A-2)
            self.out[2] = 8 / (ti.math.pi * ti.pow(self.out[0], 3))
            self.out[3] = self.m * self.out[0]

    k = K()
    k.run()
    third = F(1.0 / 3.0)
    assert k.out[0] == float(third)
    assert k.out[1] == float(F(third * F(3)) + F(0.01 * 0.1 * 0.1))
    cube = third * (third * third)                   # x**3 by squaring: x (x x)
    assert k.out[2] == float(F(8) / (F(np.pi) * cube))
    assert k.out[3] == float(F(m) * third)


def test_declared_types_stick_and_division_is_true_division():
    class K:
        f = ti.field(ti.f32, shape=6)
        n = ti.field(ti.i32, shape=3)

        @ti.kernel
        def run(self):
            cnt = 0
            acc = 0.0
            for i in range(5):
                cnt += 1
                acc += i / 2                          # i32 / int constant: f32 true division
            self.f[0] = acc
            self.n[0] = cnt
            cnt = 7.9                                 # store into an i32 local truncates
            self.n[1] = cnt
            x = 7
            self.f[1] = x / 2                         # 3.5, not 3
            self.n[2] = int(-2.5 * self.f[1])         # int() truncates toward zero
            self.f[2] = 7 % 2.5                       # constants: Python
            self.f[3] = self.f[1] % 2                 # float floor-mod: a - b floor(a / b)
            self.f[4] = ti.floor(self.f[1] / 2)
            y = -7
            self.f[5] = y % 3                         # int floor-mod: 2

    k = K()
    k.run()
    assert (k.f[0], k.n[0], k.n[1]) == (5.0, 5, 7)
    assert (k.f[1], k.n[2], k.f[2], k.f[3], k.f[4], k.f[5]) == (3.5, -8, 2.0, 1.5, 1.0, 2.0)


def test_block_scopes_declare_fresh_locals():
    class K:
        f = ti.field(ti.f32, shape=2)

        @ti.kernel
        def run(self):
            for i in range(2):
                if i == 0:
                    v = 3                             # an i32 `v` in this block ...
                    v = v / 2 + 0.75                  # ... so 2.25 is stored truncated
                    self.f[0] = v
                else:
                    v = 3.0                           # ... and a fresh f32 `v` in this one
                    v = v / 2 + 0.75
                    self.f[1] = v

    k = K()
    k.run()
    assert (k.f[0], k.f[1]) == (2.0, 2.25)


def test_template_arguments_are_references_and_plain_arguments_values():
    class K:
        f = ti.field(ti.f32, shape=3)

        @ti.func
        def each(self, i, task: ti.template(), ret: ti.template()):
            i -= 1                                    # the caller's i is untouched
            for k in range(3):
                ret += task(i, k)

        @ti.func
        def term(self, i, k):
            return i * 10 + k + 0.5

        @ti.kernel
        def run(self):
            i = 4
            s = 0.0
            self.each(i, self.term, s)                # for_all_neighbor's calling convention (ParticleSystem.py:448)
            self.f[0] = s
            self.f[1] = i
            v = ti.Vector([0.0, 0.0, 0.0])
            self.each(i, self.vterm, v)
            self.f[2] = v.dot(ti.Vector([1, 10, 100]))

        @ti.func
        def vterm(self, i, k):
            return ti.Vector([1.0, 2.0, 3.0]) * k

    k = K()
    k.run()
    assert k.f[0] == 30.5 + 31.5 + 32.5 and k.f[1] == 4.0
    assert k.f[2] == 3 * 1 + 6 * 10 + 9 * 100


def test_struct_fields_copy_on_bind_and_store_through_on_index():
    P = ti.types.struct(pos=ti.math.vec3, mass=float, index=int, cell=ti.math.ivec3)

    class K:
        p = P.field(shape=3)
        f = ti.field(ti.f32, shape=4)

        @ti.kernel
        def run(self):
            self.p[1].pos = ti.Vector([1.0, 2.0, 3.0])
            self.p.mass[1] = 0.125
            self.p.index[1] = 7
            a = self.p[1]                             # a copy of the whole struct (A-10)
            self.p[1].pos = ti.Vector([9.0, 9.0, 9.0])
            self.p.mass[1] += 1
            self.f[0] = a.pos.y
            self.f[1] = a.mass
            self.f[2] = self.p[1].pos.y
            self.f[3] = self.p.mass[1]
            self.p.pos[2][1] = 5.0                    # element store through a vector-field element
            self.p.pos[2][1] *= -0.5
            self.p.cell[2] = ti.floor(self.p[1].pos / 2.0, ti.i32)

    k = K()
    k.run()
    assert [k.f[j] for j in range(4)] == [2.0, 0.125, 9.0, 1.125]
    assert k.p.pos.to_numpy()[2].tolist() == [0.0, -2.5, 0.0]
    assert k.p.cell.to_numpy()[2].tolist() == [4, 4, 4] and k.p.index.to_numpy().tolist() == [0, 7, 0]


def test_vector_sums_run_left_to_right_without_contraction():
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal(3).astype(F), rng.standard_normal(3).astype(F)

    class K:
        v = ti.Vector.field(3, ti.f32, shape=2)
        f = ti.field(ti.f32, shape=3)

        @ti.kernel
        def run(self):
            x = self.v[0]
            y = self.v[1]
            self.f[0] = x.dot(y)
            self.f[1] = (x - y).norm()
            self.f[2] = ti.math.cross(x, y).z

    k = K()
    k.v.from_numpy(np.stack([a, b]))
    k.run()
    assert k.f[0] == float((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2])
    d = a - b
    assert k.f[1] == float(np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]))
    assert k.f[2] == float(a[0] * b[1] - a[1] * b[0])


def test_dynamic_cells_atomics_and_grouped_ranges():
    class K:
        n = ti.field(ti.i32, shape=4)
        f = ti.field(ti.f32, shape=1)

        def __init__(self):
            S = ti.root.dense(ti.i, 8).dynamic(ti.j, 512, chunk_size=32)
            self.cells = ti.field(int)
            S.place(self.cells)

        @ti.kernel
        def run(self):
            for i in range(6):
                self.cells[i % 2].append(10 + i)
            self.n[0] = self.cells[0].length()
            self.n[1] = self.cells[1, 2]              # arrival order: 11, 13, 15
            if self.cells[1].length() != 0:
                self.cells[1].deactivate()
            self.n[2] = self.cells[1].length()
            best = -1
            where = -1
            for i in range(5):
                old = ti.atomic_max(best, (i * 2) % 3)  # returns the OLD value (ParticleSystem.py:418-420)
                if old == (i * 2) % 3:
                    where = i
            self.n[3] = where * 10 + best
            lo = ti.math.inf
            order = 0.0
            for I in ti.grouped(ti.ndrange((-1, 2), (-1, 2), (-1, 2))):
                ti.atomic_min(lo, I.dot(ti.Vector([1, 3, 9])))
                if I.x == 1 and I.y == -1:
                    order = order * 10 + (I.z + 2)    # last index fastest: 1, 2, 3
            self.f[0] = lo * 1000 + order

    k = K()
    k.run()
    assert [k.n[j] for j in range(4)] == [3, 15, 0, 42]        # the old maximum equals the value at i = 4
    assert k.f[0] == -13 * 1000 + 123


def test_python_scope_sees_plain_numbers():
    class K:
        dt = ti.field(ti.f32, shape=())

        @ti.kernel
        def err(self) -> ti.f32:
            return self.dt[None] * 2

    k = K()
    k.dt[None] = 2.5e-4
    assert isinstance(k.dt[None], float) and k.dt[None] == float(F(2.5e-4))
    assert isinstance(k.err(), float) and k.err() == float(F(2.5e-4) * F(2))
    v = ti.Vector([0.7, 0.8, 0.7]) - ti.Vector([0.0, 0.0, 0.0])
    assert v.x == 0.7 and isinstance(v.x, float)      # Python scope: binary64 (compute_boundary_particles_count)
    assert int(v.x / 0.05 + 1) == 14                  # ... 14 on the host,
    assert int(F(0.7) / F(0.05) + F(1)) == 15         # ... 15 where the same expression runs on f32 locals (quirk B-18)


def test_matrices_products_inverse_and_rotation():
    rng = np.random.default_rng(5)
    A = (rng.standard_normal((3, 3)) + 3 * np.eye(3)).astype(F)
    v = rng.standard_normal(3).astype(F)

    class K:
        a = ti.field(ti.math.mat3, shape=())
        out = ti.field(ti.math.mat3, shape=2)
        w = ti.Vector.field(3, ti.f32, shape=2)

        @ti.kernel
        def run(self):
            m = self.a[None]
            self.w[0] = m @ self.w[1]
            self.out[0] = ti.math.inverse(m)
            r = ti.math.rotation3d(0.0, 0.0, 0.0)
            k_ = ti.Matrix.identity(ti.f32, 3) / 4 - m @ ti.math.mat3([[r[0, 0], r[0, 1], r[0, 2]], [r[1, 0], r[1, 1], r[1, 2]],
                                                                        [r[2, 0], r[2, 1], r[2, 2]]]).transpose()
            self.out[1] = k_

    k = K()
    k.a[None] = ti.Matrix(A.tolist())
    k.w[1] = ti.Vector(v.tolist())
    k.run()
    want = [(A[i, 0] * v[0] + A[i, 1] * v[1]) + A[i, 2] * v[2] for i in range(3)]     # left to right
    assert k.w.to_numpy()[0].tolist() == [float(x) for x in want]
    inv = k.out.to_numpy()[0]
    assert np.abs(inv.astype(np.float64) @ A.astype(np.float64) - np.eye(3)).max() < 1e-5
    det = (A[0, 0] * (A[1, 1] * A[2, 2] - A[2, 1] * A[1, 2]) - A[1, 0] * (A[0, 1] * A[2, 2] - A[2, 1] * A[0, 2])) \
        + A[2, 0] * (A[0, 1] * A[1, 2] - A[1, 1] * A[0, 2])
    assert inv[0, 0] == (F(1.0) / det) * (A[1, 1] * A[2, 2] - A[2, 1] * A[1, 2])      # 1 / det first, then the cofactor
    assert np.array_equal(k.out.to_numpy()[1], (np.eye(3, dtype=F) / F(4)) - A)        # rotation3d(0, 0, 0) is the identity
