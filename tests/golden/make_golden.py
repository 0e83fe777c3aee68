"""Generates tests/golden/oracle_small_block.json.

PROVENANCE: these vectors come from the CPU ORACLE (oracle/sph_oracle.c), not from the reference itself:
taichi==1.6.0 cannot be imported in this image (no cp312 wheel, no network) and the reference ships no
tests or golden vectors, so parity stays "unpinned".  The fixture pins the ORACLE
(a change of its arithmetic shows up as a hash mismatch) and gives the GPU tests a target that does not
need the oracle at run time.  Run from the repository root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cfd_taichi_b200 import scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402

FIELDS = {
    "dfsph": ["pos", "vel", "rho", "alpha", "warm_start_k", "nbr_count", "cell1"],
    "wcsph": ["pos", "vel", "rho", "pressure"],
    "pcisph": ["pos", "vel", "rho", "press_iter"],
    "iisph": ["pos", "vel", "rho", "p_past"],
    "pbf": ["pos", "vel", "rho", "pbf_lambda", "pbf_delta_pos"],
}
SCALARS = {"dfsph": ["df_div_iters", "df_den_iters", "delta_time"], "pcisph": ["pc_iters", "pc_delta"], "iisph": ["ii_iters"],
           "wcsph": [], "pbf": []}
STEPS = 3


def perturb(o):
    """Deterministic non-trivial state: squeeze the block a little and give it a shear velocity."""
    pos = o.field("pos").astype(np.float64)
    c = pos.mean(axis=0, keepdims=True)
    o.field("pos")[:] = (c + (pos - c) * 0.97).astype(np.float32)
    v = np.zeros_like(pos)
    v[:, 0] = 2.0 * (pos[:, 1] - c[0, 1])
    v[:, 1] = -0.5
    o.field("vel")[:] = v.astype(np.float32)


def digest(a):
    a = np.ascontiguousarray(a)
    a = np.where(a == 0, np.zeros_like(a), a) if a.dtype.kind == "f" else a     # -0.0 and +0.0 hash alike
    return hashlib.sha256(a.tobytes()).hexdigest()


def main():
    out = {"provenance": "CPU oracle (oracle/sph_oracle.c), NOT the reference: parity unpinned", "steps": STEPS,
           "scene": "small_block (the reference's *_config_backup.json scenes, N = 5879)", "solvers": {}}
    for solver, fields in FIELDS.items():
        cfg = scenes.shipped("small_block", solver)
        o = O.Oracle(cfg, solver=solver, threads=1)
        perturb(o)
        o.step(STEPS)
        rec = {"fields": {}, "scalars": {}}
        for f in fields:
            a = o.field(f)
            rec["fields"][f] = {"sha256": digest(a), "shape": list(a.shape), "head": np.asarray(a).reshape(-1)[:6].tolist(),
                                "sum": float(np.asarray(a, dtype=np.float64).sum())}
        for s in SCALARS[solver]:
            rec["scalars"][s] = float(o.scalar(s))
        out["solvers"][solver] = rec
        o.close()
    sizes = {}
    for name in ("default", "breaking_dam_30k", "dam_flush_cube", "small_block"):
        pn, bn, g = O.derived_sizes(scenes.shipped(name, "dfsph"))
        sizes[name] = {"particle_num": pn, "boundary_particles_num": bn, "grid_num": list(g)}
    for n in (100, 160, 200):
        pn, bn, g = O.derived_sizes(scenes.breaking_dam(n))
        sizes["breaking_dam_%d" % n] = {"particle_num": pn, "boundary_particles_num": bn, "grid_num": list(g)}
    out["derived_sizes"] = sizes
    with open(os.path.join(ROOT, "tests", "golden", "oracle_small_block.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote tests/golden/oracle_small_block.json")


if __name__ == "__main__":
    main()
