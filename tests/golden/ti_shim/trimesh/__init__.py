"""Stand-in for `trimesh` next to the Taichi stand-in (tests/golden/ti_shim/taichi): just the calls ParticleSystem.py:42-57
makes -- load_mesh, apply_scale, voxelized(pitch)[.fill()].points, vertices.  The mesh reader and the voxeliser are this
repository's restatement (cfd_taichi_b200/scene.py, SURVEY.md B-R1: unverifiable without trimesh); the oracle and the CUDA
path are given the same points, so the fixtures pin the solver arithmetic on them, not the voxelisation."""
import importlib.util
import os

# loaded by path: the repository root must not enter sys.path here (its root-level ParticleSystem.py / solver_base.py are
# the product's drop-in modules and would shadow the reference's files the generator is importing)
_ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", ".."))
_spec = importlib.util.spec_from_file_location("_refshim_scene", os.path.join(_ROOT, "cfd_taichi_b200", "scene.py"))
_scene = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_scene)


class _Voxels:
    def __init__(self, mesh, pitch, filled):
        self._mesh, self._pitch, self._filled = mesh, pitch, filled

    def fill(self):
        return _Voxels(self._mesh, self._pitch, True)

    @property
    def points(self):
        return _scene.voxelize(self._mesh.vertices, self._mesh.faces, self._pitch, self._filled)


class _Mesh:
    def __init__(self, vertices, faces):
        self.vertices, self.faces = vertices, faces

    def apply_scale(self, s):
        self.vertices = self.vertices * s
        return self

    def voxelized(self, pitch):
        return _Voxels(self, pitch, False)

    def export(self, *a, **k):
        raise NotImplementedError("shim: mesh export")


def load_mesh(path):
    v, f = _scene.load_mesh(path)
    return _Mesh(v, f)
