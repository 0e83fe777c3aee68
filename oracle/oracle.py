"""ctypes front-end of the CPU oracle (oracle/sph_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.  PARITY UNPINNED: see sph_oracle.h.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsph_oracle.so")

SOLVER_IDS = {"wcsph": 0, "pcisph": 1, "iisph": 2, "dfsph": 3, "pbf": 4}


class OrcConfig(ctypes.Structure):
    _fields_ = [
        ("box_min", ctypes.c_double * 3),
        ("box_max", ctypes.c_double * 3),
        ("particle_radius", ctypes.c_double),
        ("gravity", ctypes.c_double),
        ("start_pos", ctypes.c_double * 3),
        ("water_size", ctypes.c_double * 3),
        ("delta_time", ctypes.c_double),
        ("boundary_handle", ctypes.c_int),
        ("fs_couple", ctypes.c_int),
        ("solver", ctypes.c_int),
        ("exist_rigid", ctypes.c_int),
        ("active_rigid", ctypes.c_int),
        ("rigid_rho", ctypes.c_double),
        ("rigid_pos_offset", ctypes.c_double * 3),
        ("rigid_att_offset_deg", ctypes.c_double * 3),
        ("n_rigid_vertices", ctypes.c_int),
    ]


def build(force=False):
    """Compile the oracle with the committed Makefile (building the checker is not using it)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("sph_oracle.c", "sph_oracle_solvers2.inc", "sph_oracle_pbf.inc", "sph_oracle.h")
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.POINTER(OrcConfig), ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_step.argtypes = [ctypes.c_void_p]
        L.orc_rigid_step.argtypes = [ctypes.c_void_p]
        L.orc_base_step.argtypes = [ctypes.c_void_p]
        L.orc_phase.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.orc_field.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p),
                                ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_int),
                                ctypes.POINTER(ctypes.c_int)]
        L.orc_field.restype = ctypes.c_int
        L.orc_scalar.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.orc_scalar.restype = ctypes.c_double
        L.orc_set_scalar.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
        L.orc_set_threads.argtypes = [ctypes.c_int]
        L.orc_derived_sizes.argtypes = [ctypes.POINTER(OrcConfig), ctypes.POINTER(ctypes.c_longlong),
                                        ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_int)]
        L.orc_cubic_kernel.argtypes = [ctypes.c_float, ctypes.c_float]
        L.orc_cubic_kernel.restype = ctypes.c_float
        L.orc_cubic_kernel_derivative.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_float,
                                                  ctypes.POINTER(ctypes.c_float)]
        L.orc_poly_kernel.argtypes = [ctypes.c_float, ctypes.c_float]
        L.orc_poly_kernel.restype = ctypes.c_float
        L.orc_spiky_kernel_derivative.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_float,
                                                  ctypes.POINTER(ctypes.c_float)]
        L.orc_cull_threshold.argtypes = [ctypes.c_float]
        L.orc_cull_threshold.restype = ctypes.c_float
        _lib = L
    return _lib


def make_config(config, solver=None):
    """JSON dict (reference schema) -> OrcConfig."""
    scene, sol, fluid = config["scene"], config["solver"], config["fluid"]
    solid = config.get("solid", {})
    c = OrcConfig()
    for k in range(3):
        c.box_min[k] = scene["box_min"][k]
        c.box_max[k] = scene["box_max"][k]
        c.start_pos[k] = fluid["start_pos"][k]
        c.water_size[k] = fluid["water_size"][k]
    c.particle_radius = scene["particle_radius"]
    c.gravity = scene["gravity"]
    c.delta_time = sol["delta_time"]
    c.boundary_handle = 1 if sol.get("boundary_handle", True) else 0
    c.fs_couple = 1 if sol.get("fs_couple", True) else 0
    c.solver = SOLVER_IDS[solver or sol["name"]]
    c.exist_rigid = 1 if solid else 0
    if solid:
        c.active_rigid = 1 if solid.get("active", False) else 0
        c.rigid_rho = solid["rho_0"]
        for k in range(3):
            c.rigid_pos_offset[k] = solid["pos_offset"][k]
            c.rigid_att_offset_deg[k] = solid["attitude_offset"][k]
    return c


def derived_sizes(config):
    c = make_config(config)
    pn, bn = ctypes.c_longlong(), ctypes.c_longlong()
    g = (ctypes.c_int * 3)()
    lib().orc_derived_sizes(ctypes.byref(c), ctypes.byref(pn), ctypes.byref(bn), g)
    return pn.value, bn.value, (g[0], g[1], g[2])


class Oracle:
    """One simulation in the CPU oracle.  `field(name)` returns a live numpy VIEW."""

    def __init__(self, config, solver=None, rigid_points=None, rigid_vertices=None, threads=1):
        L = lib()
        L.orc_set_threads(int(threads))
        self.cfg = make_config(config, solver)
        rp = rv = None
        n_r = 0
        if rigid_points is not None:
            self._rp = np.ascontiguousarray(rigid_points, dtype=np.float32)
            rp, n_r = self._rp.ctypes.data, self._rp.shape[0]
        if rigid_vertices is not None:
            self._rv = np.ascontiguousarray(rigid_vertices, dtype=np.float32)
            rv = self._rv.ctypes.data
            self.cfg.n_rigid_vertices = self._rv.shape[0]
        self._h = L.orc_create(ctypes.byref(self.cfg), rp, n_r, rv)
        if not self._h:
            raise RuntimeError("orc_create failed")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, n=1, rigid=True):
        for _ in range(n):
            lib().orc_step(self._h)
            if rigid:
                lib().orc_rigid_step(self._h)

    def base_step(self):
        lib().orc_base_step(self._h)

    def phase(self, name):
        lib().orc_phase(self._h, name.encode())

    def field(self, name):
        p = ctypes.c_void_p()
        n = ctypes.c_longlong()
        nc = ctypes.c_int()
        isf = ctypes.c_int()
        if lib().orc_field(self._h, name.encode(), ctypes.byref(p), ctypes.byref(n), ctypes.byref(nc),
                           ctypes.byref(isf)) != 0:
            raise KeyError(name)
        cnt = n.value * nc.value
        ct = ctypes.c_float if isf.value else ctypes.c_int
        if cnt == 0:
            return np.zeros((0, nc.value) if nc.value > 1 else (0,), dtype=np.float32 if isf.value else np.int32)
        arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ct)), shape=(cnt,))
        return arr.reshape(n.value, nc.value) if nc.value > 1 else arr

    def scalar(self, name):
        return lib().orc_scalar(self._h, name.encode())

    def set_scalar(self, name, v):
        lib().orc_set_scalar(self._h, name.encode(), float(v))


def cubic_kernel(r, h):
    return lib().orc_cubic_kernel(float(r), float(h))


def cubic_kernel_derivative(r, h):
    a = (ctypes.c_float * 3)(*[float(x) for x in r])
    o = (ctypes.c_float * 3)()
    lib().orc_cubic_kernel_derivative(a, float(h), o)
    return np.array([o[0], o[1], o[2]], dtype=np.float32)


def poly_kernel(r, h):
    return lib().orc_poly_kernel(float(r), float(h))


def spiky_kernel_derivative(r, h):
    a = (ctypes.c_float * 3)(*[float(x) for x in r])
    o = (ctypes.c_float * 3)()
    lib().orc_spiky_kernel_derivative(a, float(h), o)
    return np.array([o[0], o[1], o[2]], dtype=np.float32)


def cull_threshold(h):
    return lib().orc_cull_threshold(float(h))
