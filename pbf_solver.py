# coding=utf-8
"""Top-level module name of the reference (`pbf_solver.py`), so that `main.py:10-11, 65-68` of the reference
(`from ParticleSystem import ParticleSystem`, `importlib.import_module(name + '_solver')`, `utils.read_config`)
resolve unchanged.  The implementation lives in cfd_taichi_b200/pbf_solver.py."""
from cfd_taichi_b200.pbf_solver import pbf_solver  # noqa: F401
