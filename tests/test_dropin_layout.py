"""The names a user of the reference types keep working: root-level modules, shipped scene
files with the sizes of SURVEY Appendix C, the voxeliser on the shipped box mesh.  CPU only."""
import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from cfd_taichi_b200 import scene, scenes

# file -> (solver.name, N fluid, N boundary, grid) : SURVEY Appendix C, derived with PS:83-86,100-101,129-137
APPENDIX_C = {
    "default.json": ("pcisph", 132479, 67202, (71, 71, 26)),
    "config/breaking_dam_30k.json": ("iisph", 29120, 21602, (51, 31, 16)),
    "config/breaking_dam_demo.json": ("dfsph", 156799, 96802, (101, 71, 31)),
    "config/coupling_demo.json": ("pcisph", 55200, 52002, (51, 71, 26)),
    "config/dam_flush_cube.json": ("pcisph", 56447, 21602, (51, 31, 16)),
    "config/experiment1_config.json": ("iisph", 22400, 10682, (26, 25, 16)),
    "config/experiment2_config.json": ("wcsph", 22400, 10682, (26, 25, 16)),
    "config/dfsph_config_backup.json": ("dfsph", 5879, 9002, (16, 31, 16)),
    "config/iisph_config_backup.json": ("iisph", 5879, 9002, (16, 31, 16)),
    "config/pbf_config_backup.json": ("pbf", 5879, 9002, (16, 31, 16)),
    "config/pcisph_config_backup.json": ("pcisph", 5879, 9002, (16, 31, 16)),
    "config/wcsph_config_backup.json": ("wcsph", 5879, 9002, (16, 31, 16)),
    "config/dam_1m_dfsph.json": ("dfsph", 1000000, 191682, (151, 81, 53)),
    "config/dam_4m_pcisph.json": ("pcisph", 4096000, 465122, (241, 121, 83)),
    "config/dam_4m_iisph.json": ("iisph", 4096000, 465122, (241, 121, 83)),
    "config/dam_8m_dfsph.json": ("dfsph", 8000000, 725402, (301, 151, 103)),
}


@pytest.mark.parametrize("rel", sorted(APPENDIX_C))
def test_shipped_scene_files(rel):
    import utils                      # the root-level module name of the reference (main.py:11)
    cfg = utils.read_config(os.path.join(ROOT, rel))
    name, n, nb, grid = APPENDIX_C[rel]
    assert cfg["solver"]["name"] == name
    assert scene.derive_sizes(cfg) == (n, nb, grid)


def test_scene_files_are_what_the_generator_writes(tmp_path):
    # config/make_configs.py is the provenance of every shipped scene file and of the two box meshes
    for fname, (sc, solver) in scenes.FILES.items():
        path = os.path.join(ROOT, "default.json" if fname == "default" else os.path.join("config", fname + ".json"))
        with open(path) as f:
            assert json.load(f) == json.loads(json.dumps(scenes.shipped(sc, solver))), fname


def test_root_level_module_names_resolve():
    """main.py:10-11, 65-68 of the reference: `from ParticleSystem import ParticleSystem`,
    `importlib.import_module(name + '_solver')`, `getattr(module, name + '_solver')`, `utils.read_config`."""
    code = ("import importlib, sys; sys.path.insert(0, %r)\n"
            "from ParticleSystem import ParticleSystem\n"
            "from rigid_solver import rigid_solver\n"
            "import utils, solver_base\n"
            "for n in ('wcsph', 'pcisph', 'iisph', 'dfsph', 'pbf'):\n"
            "    m = importlib.import_module(n + '_solver'); c = getattr(m, n + '_solver')\n"
            "    assert issubclass(c, solver_base.solver_base), n\n"
            "assert callable(utils.read_config) and ParticleSystem.__name__ == 'ParticleSystem'\n"
            "print('ok')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], cwd="/", capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_main_without_a_gpu_fails_loudly_not_silently():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    out = subprocess.run([sys.executable, "main.py", "--config", "config/dfsph_config_backup.json", "--steps", "1"],
                         cwd=ROOT, capture_output=True, text=True)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def _sorted_rows(a):
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def test_voxeliser_on_the_shipped_box_mesh():
    """ParticleSystem.py:42-50 on obj/cube1.STL (0.8 x 0.5 x 1.0): 17 x 11 x 21 = 3 927 voxel centres at pitch
    2 * voxel_radius, 8 mesh vertices (SURVEY B-R1)."""
    solid = scenes.shipped("dam_flush_cube")["solid"]
    pts, verts, faces = scene.rigid_points_from_config(solid, ROOT)
    assert pts.shape == (3927, 3) and verts.shape == (8, 3) and faces.shape == (12, 3)
    ax = [np.arange(n) * np.float64(0.05) for n in (17, 11, 21)]
    want = np.stack(np.meshgrid(*ax, indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    assert np.array_equal(_sorted_rows(pts), _sorted_rows(want))
    ref = "/root/reference/obj/cube1.STL"
    if os.path.exists(ref):           # the reference's own asset gives the same particles (checked where it exists)
        pts_ref, verts_ref, _ = scene.rigid_points_from_config(dict(solid, mesh=ref), ROOT)
        assert np.array_equal(pts_ref, pts) and verts_ref.shape == (8, 3)


def test_lower_case_mesh_extension_resolves():
    # several shipped files spell "./obj/cube1.stl" (SURVEY B-R5); on a case-sensitive file system the loader
    # falls back to the existing spelling instead of failing
    solid = scenes.shipped("coupling_demo")["solid"]
    assert solid["mesh"].endswith(".stl")
    pts, _, _ = scene.rigid_points_from_config(solid, ROOT)
    assert pts.shape == (3927, 3)
