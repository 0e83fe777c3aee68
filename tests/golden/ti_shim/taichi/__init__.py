"""A stand-in for the `taichi` package, just large enough to EXECUTE the reference's own solver sources
(ParticleSystem.py, solver_base.py, dfsph/wcsph/pcisph/iisph_solver.py, rigid_solver.py) on the CPU, unmodified, where Taichi itself
cannot be installed (no wheel for this interpreter, no network).  TEST INFRASTRUCTURE: used only by
tests/golden/make_reference_shim_golden.py (which writes the committed fixtures) and by the CPU test that re-runs a
small case when /root/reference is present.  Nothing in the product imports it.

How it works (the same architecture as Taichi's own front end): `@ti.kernel` / `@ti.func` read the decorated
function's source, rewrite its AST and compile the result; the rewritten function then runs as ordinary Python on
value classes that carry Taichi's arithmetic:

  * F32 / I32 are the runtime scalar types (binary32 held in a numpy float32, round-to-nearest-even on every single
    operation, no contraction; 32-bit integers).  Python floats / ints met inside a kernel are COMPILE-TIME constants,
    exactly as in Taichi: constant (x) constant is evaluated by Python in binary64, constant (x) runtime value casts the
    constant to the runtime type first.
  * assignment to a new local declares it with the value's type (float constant -> f32, int constant -> i32); assignment
    to an existing local casts to the declared type; scopes are the blocks of the source (if / for / while bodies).
  * `/` is true division (i32 / i32 -> f32), `%` and `//` follow floor semantics, `**` with an integer exponent is
    exponentiation by squaring, int() truncates, ti.floor(x, ti.i32) floors then casts (appendix A-4, A-5).
  * vectors: norm = sqrt((x x + y y) + z z), dot = (a0 b0 + a1 b1) + a2 b2, element-wise operators (A-6).
  * struct fields are array-of-structures; `field[i]` is an l-value, binding it to a local copies the struct (A-10).
  * ti.func arguments are passed by value, `ti.template()` arguments by reference (for_all_neighbor's `ret`).
  * top-level loops run in ascending order on one thread (the reference's own cpu_max_num_threads=1 order), `+=` on a
    kernel local from inside a loop is the sequential sum, ti.atomic_max / atomic_min return the OLD value (A-8).
  * dynamic SNodes are per-cell Python lists: append order = arrival order, deactivate() empties the cell (A-9).
  * matrices (rigid body): products accumulate left to right; inverse / determinant / ti.math.rotation3d are Taichi's
    closed forms as recalled (A-6, A-11: the two items of the appendix that cannot be checked offline).

What it is NOT: Taichi.  Every rule above is this repository's reading of Taichi 1.6; the
reference's own statements, loop structure, operand order, constants and quirks, however, are executed as written,
which is what the restated oracle (oracle/sph_oracle.c) has to be checked against.
"""
import ast
import builtins as _b
import inspect
import math as _pm
import textwrap
import types as _pt

import numpy as np

_f32 = np.float32
_STATE = {"depth": 0}
_OUTS = [()]
_NOREF = object()


class _DType:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "ti." + self.name


f32 = float32 = _DType("f32")
i32 = int32 = _DType("i32")
f64 = float64 = _DType("f64")


def _dt(x):
    if x is float or x is f32:
        return f32
    if x is int or x is i32:
        return i32
    if isinstance(x, (_VecType, _MatType, _DType)):
        return x
    raise TypeError("shim: unsupported dtype %r" % (x,))


def _in_kernel():
    return _STATE["depth"] > 0


# ---------------------------------------------------------------------------------------------------------------
# runtime scalars
# ---------------------------------------------------------------------------------------------------------------
def _tof(x):
    """any scalar -> numpy float32 (the cast Taichi inserts when a value meets an f32 expression)"""
    if type(x) is F32:
        return x.v
    if type(x) is I32:
        return _f32(x.i)
    return _f32(x)


def _is_rt(x):
    t = type(x)
    return t is F32 or t is I32


def _is_floaty(x):
    return type(x) is F32 or isinstance(x, (float, np.floating))


def _ipow(a, n):
    """x ** n for an integer n >= 0 by squaring, least significant bit first (appendix A-5): x**3 = x (x x),
    x**7 = (x (x x)) ((x x)(x x))"""
    if n < 0:
        raise NotImplementedError("shim: negative integer exponent")
    if n == 0:
        return _f32(1.0)
    result, p = None, a
    while n > 0:
        if n & 1:
            result = p if result is None else result * p
        n >>= 1
        if n:
            p = p * p
    return result


class F32:
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v if type(v) is _f32 else _f32(v)

    # arithmetic: one correctly rounded binary32 operation each
    def __add__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        return F32(self.v + _tof(o))

    def __radd__(self, o):
        return F32(_tof(o) + self.v)

    def __sub__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        return F32(self.v - _tof(o))

    def __rsub__(self, o):
        return F32(_tof(o) - self.v)

    def __mul__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        return F32(self.v * _tof(o))

    def __rmul__(self, o):
        return F32(_tof(o) * self.v)

    def __truediv__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        with np.errstate(all="ignore"):
            return F32(self.v / _tof(o))

    def __rtruediv__(self, o):
        with np.errstate(all="ignore"):
            return F32(_tof(o) / self.v)

    def __mod__(self, o):
        b = _tof(o)
        with np.errstate(all="ignore"):
            return F32(self.v - b * np.floor(self.v / b))

    def __rmod__(self, o):
        a = _tof(o)
        with np.errstate(all="ignore"):
            return F32(a - self.v * np.floor(a / self.v))

    def __floordiv__(self, o):
        with np.errstate(all="ignore"):
            return F32(np.floor(self.v / _tof(o)))

    def __pow__(self, o):
        if type(o) is I32:
            return F32(_ipow(self.v, o.i))
        if isinstance(o, (int, np.integer)) and not isinstance(o, bool):
            return F32(_ipow(self.v, int(o)))
        return F32(np.power(self.v, _tof(o)))

    def __rpow__(self, o):
        return F32(np.power(_tof(o), self.v))

    def __neg__(self):
        return F32(-self.v)

    def __pos__(self):
        return self

    def __abs__(self):
        return F32(np.abs(self.v))

    # comparisons are made in binary32
    def __lt__(self, o):
        return bool(self.v < _tof(o))

    def __le__(self, o):
        return bool(self.v <= _tof(o))

    def __gt__(self, o):
        return bool(self.v > _tof(o))

    def __ge__(self, o):
        return bool(self.v >= _tof(o))

    def __eq__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        return bool(self.v == _tof(o))

    def __ne__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        return bool(self.v != _tof(o))

    def __hash__(self):
        return hash(float(self.v))

    def __bool__(self):
        return bool(self.v != 0)

    def __float__(self):
        return float(self.v)

    def __int__(self):
        return int(self.v)

    def __repr__(self):
        return "F32(%r)" % float(self.v)

    def __format__(self, spec):
        return format(float(self.v), spec)


def _wrap_i(v):
    v = int(v)
    if not -(1 << 31) <= v < (1 << 31):
        v = (v + (1 << 31)) % (1 << 32) - (1 << 31)
    return v


class I32:
    __slots__ = ("i",)

    def __init__(self, i):
        self.i = _wrap_i(i)

    @staticmethod
    def _is_int(o):
        return type(o) is I32 or (isinstance(o, (int, np.integer)) and not isinstance(o, float))

    @staticmethod
    def _iv(o):
        return o.i if type(o) is I32 else int(o)

    def __add__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        if I32._is_int(o):
            return I32(self.i + I32._iv(o))
        return F32(_f32(self.i) + _tof(o))

    def __radd__(self, o):
        if I32._is_int(o):
            return I32(I32._iv(o) + self.i)
        return F32(_tof(o) + _f32(self.i))

    def __sub__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        if I32._is_int(o):
            return I32(self.i - I32._iv(o))
        return F32(_f32(self.i) - _tof(o))

    def __rsub__(self, o):
        if I32._is_int(o):
            return I32(I32._iv(o) - self.i)
        return F32(_tof(o) - _f32(self.i))

    def __mul__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        if I32._is_int(o):
            return I32(self.i * I32._iv(o))
        return F32(_f32(self.i) * _tof(o))

    def __rmul__(self, o):
        if I32._is_int(o):
            return I32(I32._iv(o) * self.i)
        return F32(_tof(o) * _f32(self.i))

    def __truediv__(self, o):                      # true division: both sides become f32 (appendix A-4)
        if isinstance(o, Vector):
            return NotImplemented
        with np.errstate(all="ignore"):
            return F32(_f32(self.i) / _tof(o))

    def __rtruediv__(self, o):
        with np.errstate(all="ignore"):
            return F32(_tof(o) / _f32(self.i))

    def __mod__(self, o):
        if I32._is_int(o):
            return I32(self.i % I32._iv(o))
        return F32(_f32(self.i)) % o

    def __rmod__(self, o):
        if I32._is_int(o):
            return I32(I32._iv(o) % self.i)
        return F32(_tof(o)) % self

    def __floordiv__(self, o):
        if I32._is_int(o):
            return I32(self.i // I32._iv(o))
        return F32(_f32(self.i)) // o

    def __rfloordiv__(self, o):
        if I32._is_int(o):
            return I32(I32._iv(o) // self.i)
        return F32(_tof(o)) // self

    def __pow__(self, o):
        if I32._is_int(o):
            return I32(self.i ** I32._iv(o))
        return F32(_f32(self.i)) ** o

    def __neg__(self):
        return I32(-self.i)

    def __pos__(self):
        return self

    def __abs__(self):
        return I32(_b.abs(self.i))

    def _cmp(self, o, op):
        if I32._is_int(o):
            return op(self.i, I32._iv(o))
        return bool(op(_f32(self.i), _tof(o)))

    def __lt__(self, o):
        return self._cmp(o, lambda a, b: a < b)

    def __le__(self, o):
        return self._cmp(o, lambda a, b: a <= b)

    def __gt__(self, o):
        return self._cmp(o, lambda a, b: a > b)

    def __ge__(self, o):
        return self._cmp(o, lambda a, b: a >= b)

    def __eq__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        return self._cmp(o, lambda a, b: a == b)

    def __ne__(self, o):
        if isinstance(o, Vector):
            return NotImplemented
        return self._cmp(o, lambda a, b: a != b)

    def __hash__(self):
        return hash(self.i)

    def __bool__(self):
        return self.i != 0

    def __index__(self):
        return self.i

    def __int__(self):
        return self.i

    def __float__(self):
        return float(self.i)

    def __repr__(self):
        return "I32(%d)" % self.i

    def __format__(self, spec):
        return format(self.i, spec)


def _rt_scalar(x):
    """constant or runtime scalar -> runtime scalar (what `expr_init` does with a Python constant)"""
    t = type(x)
    if t is F32 or t is I32:
        return x
    if isinstance(x, (bool, np.bool_)):
        return I32(int(x))
    if isinstance(x, (int, np.integer)):
        return I32(x)
    if isinstance(x, (float, np.floating)):
        return F32(x)
    raise TypeError("shim: not a scalar: %r" % (x,))


def _plain(x):
    """runtime scalar -> Python value (what Python scope sees)"""
    t = type(x)
    if t is F32:
        return float(x.v)
    if t is I32:
        return x.i
    if isinstance(x, np.floating):
        return float(x)
    if isinstance(x, np.integer):
        return int(x)
    return x


# ---------------------------------------------------------------------------------------------------------------
# vectors
# ---------------------------------------------------------------------------------------------------------------
def _coerce_entries(entries):
    """ti.Vector([...]) inside a kernel: one element type for the whole vector, f32 if any entry is a float"""
    if any(_is_floaty(e) for e in entries):
        return [e if type(e) is F32 else F32(_tof(e)) for e in entries]
    return [e if type(e) is I32 else I32(I32._iv(e)) for e in entries]


class Vector:
    __slots__ = ("e",)

    def __init__(self, arr, dt=None):
        if isinstance(arr, Vector):
            entries = list(arr.e)
        else:
            entries = [_plain(a) if isinstance(a, np.generic) else a for a in arr]
        if any(isinstance(a, (list, tuple, Vector)) for a in entries):
            raise TypeError("shim: nested lists make a ti.Matrix, not a ti.Vector")
        if _in_kernel():
            entries = _coerce_entries(entries)
        self.e = entries

    @staticmethod
    def _mk(entries):
        v = Vector.__new__(Vector)
        v.e = entries
        return v

    @staticmethod
    def field(n, dtype, shape=None):
        return _Field(_VecType(n, _dt(dtype)), shape)

    # element access
    x = property(lambda s: s.e[0], lambda s, v: s.__setitem__(0, v))
    y = property(lambda s: s.e[1], lambda s, v: s.__setitem__(1, v))
    z = property(lambda s: s.e[2], lambda s, v: s.__setitem__(2, v))
    w = property(lambda s: s.e[3], lambda s, v: s.__setitem__(3, v))

    def __len__(self):
        return len(self.e)

    def __iter__(self):
        return iter(self.e)

    def __getitem__(self, k):
        return self.e[int(k)]

    def __setitem__(self, k, v):
        old = self.e[int(k)]
        self.e[int(k)] = _cast_like(old, v) if _is_rt(old) else v

    def _zip(self, o, f, swap=False):
        if isinstance(o, Vector):
            if len(o.e) != len(self.e):
                raise ValueError("shim: vector length mismatch")
            return Vector._mk([f(b, a) if swap else f(a, b) for a, b in zip(self.e, o.e)])
        return Vector._mk([f(o, a) if swap else f(a, o) for a in self.e])

    def __add__(self, o):
        return self._zip(o, lambda a, b: a + b)

    def __radd__(self, o):
        return self._zip(o, lambda a, b: a + b, swap=True)

    def __sub__(self, o):
        return self._zip(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._zip(o, lambda a, b: a - b, swap=True)

    def __mul__(self, o):
        return self._zip(o, lambda a, b: a * b)

    def __rmul__(self, o):
        return self._zip(o, lambda a, b: a * b, swap=True)

    def __truediv__(self, o):
        return self._zip(o, lambda a, b: a / b)

    def __rtruediv__(self, o):
        return self._zip(o, lambda a, b: a / b, swap=True)

    def __mod__(self, o):
        return self._zip(o, lambda a, b: a % b)

    def __pow__(self, o):
        return self._zip(o, lambda a, b: a ** b)

    def __neg__(self):
        return Vector._mk([-a for a in self.e])

    def __pos__(self):
        return self

    def __matmul__(self, o):
        if isinstance(o, Vector):
            return self.dot(o)
        raise NotImplementedError("shim: matrix product")

    def _cmpv(self, o, f):
        return self._zip(o, lambda a, b: I32(1 if f(a, b) else 0) if _in_kernel() else int(bool(f(a, b))))

    def __lt__(self, o):
        return self._cmpv(o, lambda a, b: a < b)

    def __le__(self, o):
        return self._cmpv(o, lambda a, b: a <= b)

    def __gt__(self, o):
        return self._cmpv(o, lambda a, b: a > b)

    def __ge__(self, o):
        return self._cmpv(o, lambda a, b: a >= b)

    def __eq__(self, o):
        return self._cmpv(o, lambda a, b: a == b)

    def __ne__(self, o):
        return self._cmpv(o, lambda a, b: a != b)

    __hash__ = None

    def any(self):
        return any(bool(a) for a in self.e)

    def all(self):
        return all(bool(a) for a in self.e)

    def dot(self, o):
        a, b = self.e, o.e
        if len(a) != len(b):
            raise ValueError("shim: vector length mismatch")
        s = a[0] * b[0]
        for k in range(1, len(a)):
            s = s + a[k] * b[k]
        return s

    def norm_sqr(self):
        return self.dot(self)

    def norm(self):
        return sqrt(self.dot(self))

    def cross(self, o):
        return _cross(self, o)

    def cast(self, dt):
        return Vector._mk([cast(a, dt) for a in self.e])

    def to_numpy(self):
        return np.array([_plain(a) for a in self.e])

    def to_list(self):
        return [_plain(a) for a in self.e]

    def __repr__(self):
        return "Vector(%r)" % (self.e,)


class _VecRef(Vector):
    """`vector_field[i]`: reads like a Vector, `ref[j] = x` stores through to the field (dfsph_solver.py:244-250)"""
    __slots__ = ("_owner", "_index")

    def __setitem__(self, k, v):
        Vector.__setitem__(self, k, v)
        self._owner[self._index] = self


class _VecType:
    def __init__(self, n, dt):
        self.n, self.dt = n, dt

    def __call__(self, *a):
        if len(a) == 1 and isinstance(a[0], (list, tuple, Vector)):
            a = a[0]
        v = Vector(list(a))
        return v

    def __repr__(self):
        return "vec%d(%s)" % (self.n, self.dt.name)


class Matrix:
    """Small dense matrices (rigid_solver.py, ParticleSystem.py:198-295).  Products accumulate left to right:
    c_ij = (a_i0 b_0j + a_i1 b_1j) + a_i2 b_2j; inverse / determinant are the closed forms of Taichi's matrix.py as recalled
    (1 / det first, then cofactor products) -- like rotation3d below, unverifiable here."""
    __slots__ = ("r",)

    def __init__(self, rows):
        rows = [list(r.e) if isinstance(r, Vector) else list(r) for r in rows]
        if _in_kernel():
            flat = _coerce_entries([x for r in rows for x in r])
            m = len(rows[0])
            rows = [flat[k * m:(k + 1) * m] for k in range(len(rows))]
        self.r = rows

    @staticmethod
    def _mk(rows):
        a = Matrix.__new__(Matrix)
        a.r = rows
        return a

    @staticmethod
    def identity(dt, n):
        one, zero = (F32(1.0), F32(0.0)) if _in_kernel() else (1.0, 0.0)
        return Matrix._mk([[one if a == b else zero for b in range(n)] for a in range(n)])

    n = property(lambda s: len(s.r))
    m = property(lambda s: len(s.r[0]))

    def __getitem__(self, ij):
        return self.r[int(ij[0])][int(ij[1])]

    def __setitem__(self, ij, v):
        old = self.r[int(ij[0])][int(ij[1])]
        self.r[int(ij[0])][int(ij[1])] = _cast_like(old, v) if _is_rt(old) else v

    def _zip(self, o, f, swap=False):
        if isinstance(o, Matrix):
            return Matrix._mk([[f(b, a) if swap else f(a, b) for a, b in zip(ra, rb)] for ra, rb in zip(self.r, o.r)])
        if isinstance(o, Vector):
            raise TypeError("shim: matrix (op) vector")
        return Matrix._mk([[f(o, a) if swap else f(a, o) for a in ra] for ra in self.r])

    def __add__(self, o):
        return self._zip(o, lambda a, b: a + b)

    def __radd__(self, o):
        return self._zip(o, lambda a, b: a + b, swap=True)

    def __sub__(self, o):
        return self._zip(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return self._zip(o, lambda a, b: a - b, swap=True)

    def __mul__(self, o):
        return self._zip(o, lambda a, b: a * b)

    def __rmul__(self, o):
        return self._zip(o, lambda a, b: a * b, swap=True)

    def __truediv__(self, o):
        return self._zip(o, lambda a, b: a / b)

    def __neg__(self):
        return Matrix._mk([[-a for a in ra] for ra in self.r])

    def __matmul__(self, o):
        if isinstance(o, Vector):
            out = []
            for ra in self.r:
                acc = ra[0] * o.e[0]
                for k_ in range(1, len(ra)):
                    acc = acc + ra[k_] * o.e[k_]
                out.append(acc)
            return Vector._mk(out)
        rows = []
        for ra in self.r:
            row = []
            for c in range(o.m):
                acc = ra[0] * o.r[0][c]
                for k_ in range(1, len(ra)):
                    acc = acc + ra[k_] * o.r[k_][c]
                row.append(acc)
            rows.append(row)
        return Matrix._mk(rows)

    def transpose(self):
        return Matrix._mk([[self.r[a][b] for a in range(self.n)] for b in range(self.m)])

    def determinant(self):
        a = self.r
        if self.n != 3 or self.m != 3:
            raise NotImplementedError("shim: determinant of a %dx%d matrix" % (self.n, self.m))
        return (a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[1][0] * (a[0][1] * a[2][2] - a[2][1] * a[0][2])
                + a[2][0] * (a[0][1] * a[1][2] - a[1][1] * a[0][2]))

    def inverse(self):
        if self.n != 3 or self.m != 3:
            raise NotImplementedError("shim: inverse of a %dx%d matrix" % (self.n, self.m))
        inv_det = _init(1.0 / self.determinant())
        a = self.r

        def E(x, y):
            return a[x % 3][y % 3]
        out = [[None] * 3 for _ in range(3)]
        for i_ in range(3):
            for j_ in range(3):
                out[j_][i_] = inv_det * (E(i_ + 1, j_ + 1) * E(i_ + 2, j_ + 2) - E(i_ + 2, j_ + 1) * E(i_ + 1, j_ + 2))
        return Matrix._mk(out)

    def to_numpy(self):
        return np.array([[_plain(x) for x in ra] for ra in self.r])

    def __repr__(self):
        return "Matrix(%r)" % (self.r,)

    def __format__(self, spec):
        return repr(self)


class _MatType:
    def __init__(self, n, m, dt):
        self.n, self.m, self.dt = n, m, dt

    def __call__(self, *rows):
        if len(rows) == 1:
            rows = rows[0]
        return Matrix(list(rows))


def _rotation3d(ang_x, ang_y, ang_z):
    """ti.math.rotation3d = rot_yaw_pitch_roll(yaw = ang_z, pitch = ang_x, roll = ang_y) as recalled (appendix A-11)"""
    yaw, pitch, roll = _init(ang_z), _init(ang_x), _init(ang_y)
    ch, sh = cos(yaw), sin(yaw)
    cp, sp = cos(pitch), sin(pitch)
    cb, sb = cos(roll), sin(roll)
    z, o = F32(0.0), F32(1.0)
    return Matrix._mk([[ch * cb + sh * sp * sb, sb * cp, -sh * cb + ch * sp * sb, z],
                       [-ch * sb + sh * sp * cb, cb * cp, sb * sh + ch * sp * cb, z],
                       [sh * cp, -sp, ch * cp, z],
                       [z, z, z, o]])


def _cross(a, b):
    return Vector._mk([a.e[1] * b.e[2] - a.e[2] * b.e[1],
                       a.e[2] * b.e[0] - a.e[0] * b.e[2],
                       a.e[0] * b.e[1] - a.e[1] * b.e[0]])


# ---------------------------------------------------------------------------------------------------------------
# element-wise functions (ti.sqrt, ti.floor, ti.max ...)
# ---------------------------------------------------------------------------------------------------------------
def _unary(x, f_rt, f_py):
    if isinstance(x, Vector):
        return Vector._mk([_unary(a, f_rt, f_py) for a in x.e])
    if _is_rt(x):
        return f_rt(x)
    if _in_kernel() and isinstance(x, (float, np.floating)):
        return f_py(x)              # a constant stays a constant
    return f_py(x)


def sqrt(x):
    def rt(a):
        with np.errstate(all="ignore"):
            return F32(np.sqrt(_tof(a)))
    return _unary(x, rt, _pm.sqrt)


def cos(x):
    return _unary(x, lambda a: F32(np.cos(_tof(a))), _pm.cos)


def sin(x):
    return _unary(x, lambda a: F32(np.sin(_tof(a))), _pm.sin)


def floor(x, dtype=None):
    def rt(a):
        r = a if type(a) is I32 else F32(np.floor(a.v))
        return cast(r, dtype) if dtype is not None else r

    def py(a):
        r = float(_pm.floor(a)) if dtype is None else _pm.floor(a)
        return r
    return _unary(x, rt, py)


def ceil(x, dtype=None):
    def rt(a):
        r = a if type(a) is I32 else F32(np.ceil(a.v))
        return cast(r, dtype) if dtype is not None else r
    return _unary(x, rt, _pm.ceil)


def cast(x, dtype):
    dtype = _dt(dtype)
    if isinstance(x, Vector):
        return x.cast(dtype)
    if dtype is i32:
        if type(x) is F32:
            return I32(int(x.v))                # truncation toward zero
        if type(x) is I32:
            return x
        return I32(int(x)) if _in_kernel() else int(x)
    if dtype is f32:
        if _is_rt(x) or _in_kernel():
            return F32(_tof(x))
        return float(_f32(x))
    raise NotImplementedError("shim: cast to %r" % (dtype,))


def _abs(x):
    return _unary(x, lambda a: a.__abs__(), _b.abs)


abs = _abs          # noqa: A001  (ti.abs)


def pow(a, b):      # noqa: A001  (ti.pow)
    if isinstance(a, Vector):
        return a ** b
    if _is_rt(a) or _is_rt(b):
        return _rt_scalar(a) ** b
    return a ** b


def _minmax(args, pick_first):
    r = args[0]
    for b in args[1:]:
        if isinstance(r, Vector) or isinstance(b, Vector):
            raise NotImplementedError("shim: vector min / max")
        if _is_rt(r) or _is_rt(b):
            if type(r) is F32 or type(b) is F32 or _is_floaty(r) or _is_floaty(b):
                x, y = _tof(r), _tof(b)
                r = F32(x if pick_first(x, y) else y)
            else:
                x, y = I32._iv(r), I32._iv(b)
                r = I32(x if pick_first(x, y) else y)
        else:
            r = r if pick_first(r, b) else b
    return r


def max(*a):        # noqa: A001  (ti.max)
    return _minmax(a, lambda x, y: x >= y or y != y)


def min(*a):        # noqa: A001  (ti.min)
    return _minmax(a, lambda x, y: x <= y or y != y)


def static(x):
    return x


def template():
    return _TEMPLATE


_TEMPLATE = object()


def ndrange(*dims):
    return ("ndrange", dims)


def grouped(r):
    if not (isinstance(r, tuple) and r and r[0] == "ndrange"):
        raise NotImplementedError("shim: ti.grouped over %r" % (r,))
    spans = []
    for d in r[1]:
        if isinstance(d, tuple):
            spans.append((int(d[0]), int(d[1])))
        else:
            spans.append((0, int(d)))

    def gen(k, prefix):
        if k == len(spans):
            yield Vector._mk([I32(p) for p in prefix])
            return
        for v in range(spans[k][0], spans[k][1]):
            yield from gen(k + 1, prefix + [v])
    return gen(0, [])           # last index fastest (ParticleSystem.py:452: dz fastest)


def _atomic(kind, old, val):
    """returns (new value with the type of `old`, old value)"""
    if kind == "max":
        new = max(old, val)
    elif kind == "min":
        new = min(old, val)
    elif kind == "add":
        new = old + val
    elif kind == "sub":
        new = old - val
    else:
        raise NotImplementedError(kind)
    return _cast_like(old, new), old


def atomic_max(*a):
    raise RuntimeError("shim: ti.atomic_max outside a rewritten kernel")


atomic_min = atomic_add = atomic_sub = atomic_max


# ---------------------------------------------------------------------------------------------------------------
# declaration / store rules (what the AST rewrite calls)
# ---------------------------------------------------------------------------------------------------------------
def _init(v):
    """`name = value` for a NEW local: the local gets the value's runtime type, structs and vectors are copied"""
    t = type(v)
    if t is F32 or t is I32:
        return v
    if t is _StructRef:
        return v._snapshot()
    if t is _StructVal:
        return v._copy()
    if isinstance(v, Vector):
        return Vector._mk(_coerce_entries(list(v.e)))
    if isinstance(v, Matrix):
        return Matrix._mk([_coerce_entries(list(ra)) for ra in v.r])
    if isinstance(v, (bool, np.bool_, int, float, np.integer, np.floating)):
        return _rt_scalar(v)
    return v           # bound methods, fields, None ...


def _cast_like(old, new):
    to = type(old)
    if to is F32:
        return new if type(new) is F32 else F32(_tof(new))
    if to is I32:
        if type(new) is I32:
            return new
        if type(new) is F32 or isinstance(new, (float, np.floating)):
            return I32(int(_tof(new)))          # float -> int store truncates
        return I32(I32._iv(new))
    return _init(new)


def _store(old, new):
    """`name = value` for an EXISTING local: cast to the declared type"""
    if isinstance(old, Vector) and not isinstance(old, _StructVal):
        if not isinstance(new, Vector) or len(new.e) != len(old.e):
            raise TypeError("shim: vector store of %r into %r" % (new, old))
        return Vector._mk([_cast_like(o, n) for o, n in zip(old.e, new.e)])
    if isinstance(old, Matrix):
        return Matrix._mk([[_cast_like(o, n) for o, n in zip(ro, rn)] for ro, rn in zip(old.r, new.r)])
    if _is_rt(old):
        if isinstance(new, Vector):
            raise TypeError("shim: vector stored into a scalar local")
        return _cast_like(old, new)
    return _init(new)


def _int(x):
    if type(x) is F32:
        return I32(int(x.v))
    if type(x) is I32:
        return x
    return int(x)


def _float(x):
    if _is_rt(x):
        return F32(_tof(x))
    return float(x)


def _range(*a):
    return (I32(k) for k in range(*[int(v) for v in a]))


def _ret(v):
    return _init(v)


def _clear_outs():
    _OUTS[0] = ()


def _set_outs(t):
    _OUTS[0] = t


def _take(k, cur):
    o = _OUTS[0]
    if k < len(o) and o[k] is not _NOREF:
        return o[k]
    return cur


_HELPERS = {"__ti_init": _init, "__ti_store": _store, "__ti_int": _int, "__ti_float": _float, "__ti_range": _range,
            "__ti_ret": _ret, "__ti_clear_outs": _clear_outs, "__ti_set_outs": _set_outs, "__ti_take": _take,
            "__ti_atomic": _atomic, "__ti_max": max, "__ti_min": min, "__ti_abs": _abs, "__ti_NOREF": _NOREF}


# ---------------------------------------------------------------------------------------------------------------
# the AST rewrite
# ---------------------------------------------------------------------------------------------------------------
def _name(id_, ctx=None):
    return ast.Name(id=id_, ctx=ctx or ast.Load())


def _call(fn, *args):
    return ast.Call(func=_name(fn), args=list(args), keywords=[])


def _is_ti_attr(node, names):
    return (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name)
            and node.func.value.id == "ti" and node.func.attr in names)


_ATOMICS = {"atomic_max": "max", "atomic_min": "min", "atomic_add": "add", "atomic_sub": "sub"}
_BUILTINS = {"int": "__ti_int", "float": "__ti_float", "max": "__ti_max", "min": "__ti_min", "abs": "__ti_abs"}


class _Rewrite(ast.NodeTransformer):
    def __init__(self, params):
        self.scopes = [set(params)]
        self.tmp = 0

    def declared(self, n):
        return any(n in s for s in self.scopes)

    def block(self, stmts, names=()):
        self.scopes.append(set(names))
        out = []
        for s in stmts:
            r = self.visit(s)
            if isinstance(r, list):
                out.extend(r)
            elif r is not None:
                out.append(r)
        self.scopes.pop()
        return out or [ast.Pass()]

    # ---- expressions
    def visit_Call(self, node):
        self.generic_visit(node)
        if isinstance(node.func, ast.Name) and node.func.id in _BUILTINS:
            node.func = _name(_BUILTINS[node.func.id])
        return node

    # ---- statements
    def _atomic_stmts(self, call, result_target):
        kind = _ATOMICS[call.func.attr]
        dst, val = call.args[0], self.visit(call.args[1])
        self.tmp += 1
        t = "__ti_t%d" % self.tmp
        load = ast.Name(id=dst.id, ctx=ast.Load()) if isinstance(dst, ast.Name) else self.visit(_reload(dst))
        store = ast.Name(id=dst.id, ctx=ast.Store()) if isinstance(dst, ast.Name) else _restore(self.visit(_reload(dst)))
        out = [ast.Assign(targets=[_name(t, ast.Store())], value=_call("__ti_atomic", ast.Constant(kind), load, val)),
               ast.Assign(targets=[store], value=ast.Subscript(value=_name(t), slice=ast.Constant(0), ctx=ast.Load()))]
        if result_target is not None:
            old = ast.Subscript(value=_name(t), slice=ast.Constant(1), ctx=ast.Load())
            out.extend(self._assign_name(result_target, old))
        return out

    def _assign_name(self, target, value):
        n = target.id
        if self.declared(n):
            value = _call("__ti_store", _name(n), value)
        else:
            self.scopes[-1].add(n)
            value = _call("__ti_init", value)
        return [ast.Assign(targets=[ast.Name(id=n, ctx=ast.Store())], value=value)]

    def visit_Assign(self, node):
        if len(node.targets) != 1:
            raise NotImplementedError("shim: chained assignment")
        tgt = node.targets[0]
        if _is_ti_attr(node.value, _ATOMICS) and isinstance(tgt, ast.Name):
            return self._atomic_stmts(node.value, tgt)
        value = self.visit(node.value)
        if isinstance(tgt, ast.Name):
            return self._assign_name(tgt, value)
        if isinstance(tgt, (ast.Tuple, ast.List)):
            raise NotImplementedError("shim: tuple assignment in a kernel")
        return ast.Assign(targets=[self.visit(tgt)], value=value)

    def visit_AugAssign(self, node):
        value = self.visit(node.value)
        if isinstance(node.target, ast.Name):
            n = node.target.id
            if not self.declared(n):
                raise NameError("shim: augmented assignment to undeclared local %r" % n)
            return ast.Assign(targets=[ast.Name(id=n, ctx=ast.Store())],
                              value=_call("__ti_store", _name(n), ast.BinOp(left=_name(n), op=node.op, right=value)))
        return ast.AugAssign(target=self.visit(node.target), op=node.op, value=value)

    def visit_Expr(self, node):
        c = node.value
        if _is_ti_attr(c, _ATOMICS):
            return self._atomic_stmts(c, None)
        if isinstance(c, ast.Call) and not _is_ti_attr(c, {"static"}):
            names = [(k, a.id) for k, a in enumerate(c.args) if isinstance(a, ast.Name) and self.declared(a.id)]
            call = ast.Expr(value=self.visit(c))
            if not names:
                return call
            out = [ast.Expr(value=_call("__ti_clear_outs")), call]
            for k, n in names:
                out.append(ast.Assign(targets=[ast.Name(id=n, ctx=ast.Store())],
                                      value=_call("__ti_take", ast.Constant(k), _name(n))))
            return out
        return self.generic_visit(node)

    def visit_Return(self, node):
        if node.value is None:
            return node
        return ast.Return(value=_call("__ti_ret", self.visit(node.value)))

    def visit_For(self, node):
        it = node.iter
        if isinstance(it, ast.Call) and isinstance(it.func, ast.Name) and it.func.id == "range":
            it = ast.Call(func=_name("__ti_range"), args=[self.visit(a) for a in it.args], keywords=[])
        elif _is_ti_attr(it, {"static"}):
            it = it.args[0]                 # compile-time loop: plain Python ints
        else:
            it = self.visit(it)
        if not isinstance(node.target, ast.Name):
            raise NotImplementedError("shim: loop target")
        body = self.block(node.body, [node.target.id])
        return ast.For(target=node.target, iter=it, body=body, orelse=[])

    def visit_While(self, node):
        return ast.While(test=self.visit(node.test), body=self.block(node.body), orelse=[])

    def visit_If(self, node):
        test = self.visit(node.test)
        body = self.block(node.body)
        orelse = self.block(node.orelse) if node.orelse else []
        return ast.If(test=test, body=body, orelse=orelse)


def _reload(node):
    return ast.parse(ast.unparse(node), mode="eval").body


def _restore(node):
    node.ctx = ast.Store()
    return node


def _compile(fn, is_kernel):
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    fdef = tree.body[0]
    if not isinstance(fdef, ast.FunctionDef):
        raise TypeError("shim: expected a function")
    fdef.decorator_list = []
    fdef.returns = None
    params = [a.arg for a in fdef.args.args]
    by_ref = set()
    for a in fdef.args.args:
        ann = a.annotation
        if ann is not None and isinstance(ann, ast.Call) and isinstance(ann.func, ast.Attribute) and ann.func.attr == "template":
            by_ref.add(a.arg)
        a.annotation = None
    rw = _Rewrite(params)
    body = rw.block(fdef.body)
    value_params = [p for p in params if p != "self" and p not in by_ref]
    prologue = [ast.Assign(targets=[ast.Name(id=p, ctx=ast.Store())], value=_call("__ti_init", _name(p))) for p in value_params]
    visible = [p for p in params if p != "self"]
    outs = ast.Tuple(elts=[_name(p) if p in by_ref else _name("__ti_NOREF") for p in visible], ctx=ast.Load())
    # always report: a callee without by-reference parameters must not leave an inner call's report behind
    body = [ast.Try(body=body, handlers=[], orelse=[], finalbody=[ast.Expr(value=_call("__ti_set_outs", outs))])]
    fdef.body = prologue + body
    ast.fix_missing_locations(tree)
    ast.increment_lineno(tree, fn.__code__.co_firstlineno - 1)
    g = fn.__globals__
    g.update(_HELPERS)
    ns = {}
    exec(compile(tree, fn.__code__.co_filename, "exec"), g, ns)
    new = ns[fdef.name]
    new.__qualname__ = fn.__qualname__
    new.__ti_source__ = ast.unparse(tree)
    new.__ti_original__ = fn
    return new


def func(fn):
    return _compile(fn, False)


def kernel(fn):
    compiled = _compile(fn, True)

    def launch(*a, **k):
        _STATE["depth"] += 1
        try:
            r = compiled(*a, **k)
        finally:
            _STATE["depth"] -= 1
        if isinstance(r, Vector):
            return Vector._mk([_plain(x) for x in r.e])
        return _plain(r)
    launch.__name__ = fn.__name__
    launch.__qualname__ = fn.__qualname__
    launch.__ti_compiled__ = compiled
    return launch


def data_oriented(cls):
    return cls


def init(*a, **k):
    return None


cpu = "cpu"
gpu = "gpu"
i = "axis_i"
j = "axis_j"
k = "axis_k"


# ---------------------------------------------------------------------------------------------------------------
# fields
# ---------------------------------------------------------------------------------------------------------------
def _np_dt(dt):
    return np.float32 if dt is f32 else np.int32


def _load_scalar(dt, raw):
    if _in_kernel():
        return F32(raw) if dt is f32 else I32(raw)
    return float(raw) if dt is f32 else int(raw)


def _store_scalar(dt, v):
    if dt is f32:
        return _tof(v)
    if type(v) is F32 or isinstance(v, (float, np.floating)):
        return np.int32(int(_tof(v)))
    return np.int32(I32._iv(v))


class _DynCell:
    __slots__ = ("lst",)

    def __init__(self, lst):
        self.lst = lst

    def length(self):
        return I32(len(self.lst)) if _in_kernel() else len(self.lst)

    def append(self, v):
        if len(self.lst) >= 512:
            raise OverflowError("shim: dynamic SNode capacity")
        self.lst.append(I32._iv(v))

    def deactivate(self):
        del self.lst[:]


class _Field:
    """ti.field(dtype, shape) / ti.Vector.field(n, dtype, shape) / a dynamic SNode's placed field"""

    def __init__(self, dtype, shape):
        self.dtype = dtype
        self.lists = None
        self.shape = None
        self.a = None
        if shape is not None:
            self._alloc(shape)

    def _alloc(self, shape):
        if isinstance(shape, (int, np.integer)) or type(shape) is I32:
            shape = (int(shape),)
        self.shape = tuple(int(s) for s in shape)
        if len(self.shape) > 1:
            raise NotImplementedError("shim: multi-dimensional fields")
        n = self.shape[0] if self.shape else 1
        if isinstance(self.dtype, _MatType):
            self.a = np.zeros((n, self.dtype.n, self.dtype.m), dtype=_np_dt(self.dtype.dt))
        elif isinstance(self.dtype, _VecType):
            self.a = np.zeros((n, self.dtype.n), dtype=_np_dt(self.dtype.dt))
        else:
            self.a = np.zeros((n,), dtype=_np_dt(self.dtype))

    @staticmethod
    def _ix(i):
        return 0 if i is None else int(i)

    def __getitem__(self, i):
        if self.lists is not None:
            if isinstance(i, tuple):
                v = self.lists[int(i[0])][int(i[1])]
                return I32(v) if _in_kernel() else v
            return _DynCell(self.lists[int(i)])
        i = self._ix(i)
        if isinstance(self.dtype, _MatType):
            mk = (F32 if self.dtype.dt is f32 else I32) if _in_kernel() else (float if self.dtype.dt is f32 else int)
            return Matrix._mk([[mk(x) for x in ra] for ra in self.a[i]])
        if isinstance(self.dtype, _VecType):
            row = self.a[i]
            dt = self.dtype.dt
            if _in_kernel():
                ref = _VecRef.__new__(_VecRef)
                ref.e = [F32(x) for x in row] if dt is f32 else [I32(x) for x in row]
                ref._owner, ref._index = self, i
                return ref
            return Vector._mk([float(x) for x in row] if dt is f32 else [int(x) for x in row])
        return _load_scalar(self.dtype, self.a[i])

    def __setitem__(self, i, v):
        i = self._ix(i)
        if isinstance(self.dtype, _MatType):
            for a_, ra in enumerate(v.r):
                for b_, e in enumerate(ra):
                    self.a[i, a_, b_] = _store_scalar(self.dtype.dt, e)
            return
        if isinstance(self.dtype, _VecType):
            if not isinstance(v, Vector) or len(v.e) != self.dtype.n:
                raise TypeError("shim: vector field store of %r" % (v,))
            dt = self.dtype.dt
            for k_, e in enumerate(v.e):
                self.a[i, k_] = _store_scalar(dt, e)
        else:
            self.a[i] = _store_scalar(self.dtype, v)

    def fill(self, v):
        if isinstance(self.dtype, _VecType):
            if isinstance(v, Vector):
                for k_, e in enumerate(v.e):
                    self.a[:, k_] = _store_scalar(self.dtype.dt, e)
            else:
                self.a[:, :] = _store_scalar(self.dtype.dt, v)
        else:
            self.a[:] = _store_scalar(self.dtype, v)

    def from_numpy(self, arr):
        self.a[...] = np.asarray(arr).astype(self.a.dtype).reshape(self.a.shape)

    def to_numpy(self):
        out = self.a.copy()
        return out.reshape(()) if self.shape == () and not isinstance(self.dtype, (_VecType, _MatType)) else out


def field(dtype, shape=None):
    return _Field(_dt(dtype), shape)


class _SNode:
    def __init__(self, n=None, dynamic=False):
        self.n, self.is_dynamic = n, dynamic

    def dense(self, axis, n):
        return _SNode(int(n))

    def dynamic(self, axis, cap, chunk_size=None):
        return _SNode(self.n, True)

    def place(self, f):
        if not self.is_dynamic:
            raise NotImplementedError("shim: dense place")
        f.lists = [[] for _ in range(self.n)]
        f.shape = (self.n,)


root = _SNode()


# ---- struct fields (array of structures: one float row and one int row per element) ----------------------------
class _StructType:
    def __init__(self, **members):
        self.members = {}
        nf = ni = 0
        self.nested = any(isinstance(t, _StructType) for t in members.values())      # ParticlesBlocks: never instantiated
        for name, t in members.items():
            if isinstance(t, _StructType):
                continue
            t = _dt(t)
            n = t.n if isinstance(t, _VecType) else 1
            base = t.dt if isinstance(t, _VecType) else t
            if base is f32:
                self.members[name] = ("f", nf, n, isinstance(t, _VecType))
                nf += n
            else:
                self.members[name] = ("i", ni, n, isinstance(t, _VecType))
                ni += n
        self.nf, self.ni = nf, ni

    def field(self, shape):
        if self.nested:
            raise NotImplementedError("shim: nested struct fields")
        return _StructField(self, shape)


def _member_load(m, frow, irow, kernel):
    kind, off, n, is_vec = m
    row = frow if kind == "f" else irow
    if kernel:
        mk = F32 if kind == "f" else I32
        return Vector._mk([mk(row[off + k_]) for k_ in range(n)]) if is_vec else mk(row[off])
    conv = float if kind == "f" else int
    return Vector._mk([conv(row[off + k_]) for k_ in range(n)]) if is_vec else conv(row[off])


def _member_store(m, frow, irow, v):
    kind, off, n, is_vec = m
    row = frow if kind == "f" else irow
    dt = f32 if kind == "f" else i32
    if is_vec:
        if not isinstance(v, Vector) or len(v.e) != n:
            raise TypeError("shim: struct member store of %r" % (v,))
        for k_, e in enumerate(v.e):
            row[off + k_] = _store_scalar(dt, e)
    else:
        row[off] = _store_scalar(dt, v)


class _StructVal:
    """a struct held by value in a kernel local (`particle = self.fluid_particles[i]`)"""
    __slots__ = ("_t", "_f", "_i")

    def __init__(self, t, frow, irow):
        object.__setattr__(self, "_t", t)
        object.__setattr__(self, "_f", frow)
        object.__setattr__(self, "_i", irow)

    def _copy(self):
        return _StructVal(self._t, self._f.copy(), self._i.copy())

    def __getattr__(self, name):
        return _member_load(self._t.members[name], self._f, self._i, True)

    def __setattr__(self, name, v):
        _member_store(self._t.members[name], self._f, self._i, v)


class _StructRef:
    """`struct_field[i]`: an l-value"""
    __slots__ = ("_sf", "_k")

    def __init__(self, sf, k_):
        object.__setattr__(self, "_sf", sf)
        object.__setattr__(self, "_k", k_)

    def _snapshot(self):
        return _StructVal(self._sf.t, self._sf.f[self._k].copy(), self._sf.i[self._k].copy())

    def __getattr__(self, name):
        sf = self._sf
        return _member_load(sf.t.members[name], sf.f[self._k], sf.i[self._k], _in_kernel())

    def __setattr__(self, name, v):
        sf = self._sf
        _member_store(sf.t.members[name], sf.f[self._k], sf.i[self._k], v)


class _MemberField:
    """`struct_field.member`: indexable, fill / from_numpy / to_numpy"""

    def __init__(self, sf, name):
        self.sf, self.name, self.m = sf, name, sf.t.members[name]

    def _cols(self):
        kind, off, n, is_vec = self.m
        arr = self.sf.f if kind == "f" else self.sf.i
        return arr, off, n, is_vec

    def __getitem__(self, i):
        i = int(i)
        v = _member_load(self.m, self.sf.f[i], self.sf.i[i], _in_kernel())
        if isinstance(v, Vector) and _in_kernel():
            ref = _VecRef.__new__(_VecRef)
            ref.e, ref._owner, ref._index = v.e, self, i
            return ref
        return v

    def __setitem__(self, i, v):
        i = int(i)
        _member_store(self.m, self.sf.f[i], self.sf.i[i], v)

    def fill(self, v):
        arr, off, n, is_vec = self._cols()
        dt = f32 if self.m[0] == "f" else i32
        if is_vec and isinstance(v, Vector):
            for k_, e in enumerate(v.e):
                arr[:, off + k_] = _store_scalar(dt, e)
        else:
            arr[:, off:off + n] = _store_scalar(dt, v)

    def from_numpy(self, a):
        arr, off, n, is_vec = self._cols()
        arr[:, off:off + n] = np.asarray(a).astype(arr.dtype).reshape(arr.shape[0], n)

    def to_numpy(self):
        arr, off, n, is_vec = self._cols()
        out = arr[:, off:off + n].copy()
        return out if is_vec else out[:, 0]


class _StructField:
    def __init__(self, t, shape):
        n = int(shape if not isinstance(shape, (tuple, list)) else shape[0])
        self.t, self.n, self.shape = t, n, (n,)
        self.f = np.zeros((n, t.nf), dtype=np.float32)
        self.i = np.zeros((n, t.ni), dtype=np.int32)
        self._members = {name: _MemberField(self, name) for name in t.members}

    def __getattr__(self, name):
        try:
            return self.__dict__["_members"][name]
        except KeyError:
            raise AttributeError(name)

    def __getitem__(self, i):
        return _StructRef(self, int(i))


# ---------------------------------------------------------------------------------------------------------------
# ti.types / ti.math namespaces
# ---------------------------------------------------------------------------------------------------------------
def _isnan(x):
    return _unary(x, lambda a: I32(int(np.isnan(_tof(a)))), _pm.isnan)


def _isinf(x):
    return _unary(x, lambda a: I32(int(np.isinf(_tof(a)))), _pm.isinf)


types = _pt.SimpleNamespace(struct=lambda **m: _StructType(**m), vector=lambda n=3, dt=f32: _VecType(n, _dt(dt)))
math = _pt.SimpleNamespace(pi=_pm.pi, inf=_pm.inf, vec3=_VecType(3, f32), vec4=_VecType(4, f32), vec2=_VecType(2, f32),
                           ivec3=_VecType(3, i32), mat3=_MatType(3, 3, f32), mat4=_MatType(4, 4, f32), cross=_cross,
                           rotation3d=_rotation3d, inverse=lambda a: a.inverse(), cos=cos, sin=sin,
                           isnan=_isnan, isinf=_isinf, sqrt=sqrt, floor=floor, pow=pow, max=max, min=min)
