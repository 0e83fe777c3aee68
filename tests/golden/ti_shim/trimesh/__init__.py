"""Stand-in for `trimesh` next to the Taichi stand-in (tests/golden/ti_shim/taichi): the reference imports it at module
level (ParticleSystem.py:3) but only uses it for scenes with a rigid body, which the stand-in does not run."""


def load_mesh(*a, **k):
    raise NotImplementedError("shim: scenes with a rigid body are not supported")
