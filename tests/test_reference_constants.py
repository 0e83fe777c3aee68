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
