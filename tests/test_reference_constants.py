"""Reference-held constants.  The reference ships no test vectors, but its solver constructors hold every tuning
constant the algorithms use (iteration caps, thresholds, viscosity / tension coefficients, restitution).  Where the
checkout exists (this container), parse them out of the reference's source and hold the host mirrors AND the oracle
/ CUDA sources to the same literals.  CPU only; skipped on the GPU box (no /root/reference there)."""
import os
import re

import pytest

from conftest import ROOT

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
ASSIGN = re.compile(r"^\s*self\.(\w+)\s*=\s*([-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?)\s*(?:#.*)?$")


def literals(path):
    out = {}
    for line in open(path):
        m = ASSIGN.match(line)
        if m:
            out[m.group(1)] = float(m.group(2))
    return out


@pytest.mark.parametrize("name", ["solver_base", "dfsph_solver", "wcsph_solver", "pcisph_solver", "iisph_solver"])
def test_mirror_constructors_hold_the_references_constants(name):
    ref = literals(os.path.join(REF, name + ".py"))
    mine = literals(os.path.join(ROOT, "cfd_taichi_b200", name + ".py"))
    assert ref, name
    shared = sorted(set(ref) & set(mine))
    assert len(shared) >= 3, (name, shared)
    for k in shared:
        assert mine[k] == ref[k], "%s.%s: mirror %r, reference %r" % (name, k, mine[k], ref[k])
    # every numeric constant of the reference constructor is mirrored (same attribute name)
    missing = sorted(set(ref) - set(mine))
    assert not missing, "%s: constants of the reference without a mirror attribute: %s" % (name, missing)


def test_native_sources_use_the_references_loop_constants():
    """The literals the device-side loop control and the oracle are built on, against the reference's lines."""
    df = open(os.path.join(REF, "dfsph_solver.py")).read()
    pc = open(os.path.join(REF, "pcisph_solver.py")).read()
    ii = open(os.path.join(REF, "iisph_solver.py")).read()
    ctl = open(os.path.join(ROOT, "cfd_taichi_b200", "csrc", "sph_ctl.cuh")).read()
    orc = open(os.path.join(ROOT, "oracle", "sph_oracle.c")).read() + open(os.path.join(ROOT, "oracle", "sph_oracle_solvers2.inc")).read()
    # DFSPH: 15 divergence passes, threshold 10, break at 1e-5, at least 2 density passes, 0.1 % of rho_0
    assert "self.max_iteration_density_divergence = 15" in df and "it < 15" in ctl
    assert "self.density_divergence_threshold = 10" in df and "avg > 10.0f" in ctl
    assert "ti.abs(rho_divergence_avg - past_rho_divergence_avg) < 1e-5" in df and "< 1e-5" in ctl
    assert "self.min_iteration_density = 2" in df and "it < 2" in open(os.path.join(ROOT, "cfd_taichi_b200", "csrc", "sph_sweeps.cu")).read()
    assert "self.density_threshold = 0.1" in df
    # PCISPH: at most 80 passes, 0.1 % error; IISPH: at most 180 passes, omega 0.5
    assert "self.max_iteration = 80" in pc and "it < 80" in ctl
    assert "self.max_iter_cnt = 180" in ii and "l < 180" in ctl
    assert "self.omega = 0.5" in ii
    for lit in ("0.9999", "15", "80", "180"):
        assert lit in orc


# ---- the reference's own HOST code, executed ------------------------------------------------------------------
# ParticleSystem.__init__ derives the particle count, the boundary particle count and the grid size in Python scope
# (PS:78-102, 129-137) before any Taichi kernel runs.  Those statements are extracted from the reference's source with
# `ast`, executed unmodified against a minimal stand-in for the Python-scope `ti.Vector` (a list of Python floats with
# .x/.y/.z, -, / and to_numpy) and compared with scene.derive_sizes -- the numbers every array of the CUDA path is
# sized with -- for every shipped scene and every synthetic BASELINE block.  This IS reference output, for the sizes.
class _Vec:
    def __init__(self, v):
        self.v = list(v)
    x = property(lambda s: s.v[0]); y = property(lambda s: s.v[1]); z = property(lambda s: s.v[2])
    def __sub__(self, o): return _Vec([a - b for a, b in zip(self.v, o.v)])
    def __truediv__(self, k): return _Vec([a / k for a in self.v])
    def __mul__(self, k): return _Vec([a * k for a in self.v])
    def __getitem__(self, i): return self.v[i]
    def to_numpy(self):
        import numpy as np
        return np.array(self.v)


def _reference_sizes(config):
    import ast
    import math
    import types
    import numpy as np
    tree = ast.parse(open(os.path.join(REF, "ParticleSystem.py")).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "ParticleSystem"][0]
    init = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__init__"][0]
    count = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "compute_boundary_particles_count"][0]
    want = {"water_size", "start_pos", "particle_radius", "particle_diameter", "support_radius", "particle_m", "particle_num",
            "box_max", "box_min", "boundary_particles_num", "grid_num"}
    keep = []
    for st in init.body:
        if isinstance(st, ast.Assign) and len(st.targets) == 1:
            t = st.targets[0]
            if isinstance(t, ast.Attribute) and isinstance(t.value, ast.Name) and t.value.id == "self" and t.attr in want:
                keep.append(st)
            elif isinstance(t, ast.Name) and t.id in ("scene_config", "solver_config", "fluid_config", "solid_config", "grid_num_np"):
                keep.append(st)
    assert {s.targets[0].attr for s in keep if isinstance(s.targets[0], ast.Attribute)} == want
    fn = ast.FunctionDef(name="derive", args=init.args, body=keep, decorator_list=[], returns=None, type_comment=None, type_params=[])
    count.decorator_list = []
    mod = ast.Module(body=[count, fn], type_ignores=[])
    ast.fix_missing_locations(mod)
    ti = types.SimpleNamespace(Vector=_Vec, ceil=math.ceil, math=types.SimpleNamespace(pi=math.pi))
    ns = {"ti": ti, "np": np}
    exec(compile(mod, "<reference ParticleSystem.__init__ (sizes)>", "exec"), ns)
    self = types.SimpleNamespace()
    self.compute_boundary_particles_count = lambda: ns["compute_boundary_particles_count"](self)
    ns["derive"](self, config)
    return self.particle_num, self.boundary_particles_num, tuple(int(g) for g in self.grid_num.v), self.particle_m


def test_derived_sizes_are_the_references_own_host_code():
    import json
    from cfd_taichi_b200 import scene, scenes
    cases = []
    for dirpath in (REF, os.path.join(REF, "config")):
        for f in sorted(os.listdir(dirpath)):
            if f.endswith(".json"):
                with open(os.path.join(dirpath, f)) as fh:
                    cases.append((f, json.load(fh)))             # the reference's own scene files
    for n_side, gx in ((100, 1), (160, 1), (200, 1), (200, 8)):      # BASELINE.json's synthetic blocks
        cases.append(("synthetic %d^3 x %d" % (n_side, gx), scenes.breaking_dam(n_side, gpus_x=gx)))
    assert len(cases) >= 12
    for name, cfg in cases:
        pn, bn, grid, m = _reference_sizes(cfg)
        assert scene.derive_sizes(cfg) == (pn, bn, grid), name
        assert m == 1000 * (cfg["scene"]["particle_radius"] ** 3) * 8
    # and the shipped copies of those scene files describe the same scenes
    for f in ("default.json",) + tuple("config/" + x for x in sorted(os.listdir(os.path.join(REF, "config"))) if x.endswith(".json")):
        mine = os.path.join(ROOT, f)
        if os.path.exists(mine):
            a, b = json.load(open(mine)), json.load(open(os.path.join(REF, f)))
            assert scene.derive_sizes(a) == scene.derive_sizes(b), f
            assert a["solver"]["name"] == b["solver"]["name"] and a["solver"]["delta_time"] == b["solver"]["delta_time"], f
