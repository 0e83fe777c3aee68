"""CPU tests of the host-side mirror and of the C-ABI surface (no compute calls: no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from cfd_taichi_b200 import _lib, scene, scenes
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["default", "breaking_dam_30k", "dam_flush_cube", "small_block"])
def test_host_init_matches_oracle_bit_exact(name):
    cfg = scenes.shipped(name, "wcsph")
    cfg.pop("solid", None)
    pn, bn, g = scene.derive_sizes(cfg)
    assert (pn, bn, g) == O.derived_sizes(cfg)
    o = O.Oracle(cfg, solver="wcsph")
    assert np.array_equal(scene.init_fluid_positions(cfg, pn), o.field("pos"))
    assert np.array_equal(scene.init_boundary_positions(cfg, bn), o.field("bpos"))
    o.close()


def test_particle_num_truncation_quirk():
    # SURVEY B-1: fp64 round-off drops the last lattice site for water 0.7 x 1.5 x 0.7 and 1.8 x 2.8 x 1.4
    assert scene.derive_sizes(scenes.shipped("small_block"))[0] == 14 * 30 * 14 - 1
    assert scene.derive_sizes(scenes.shipped("dam_flush_cube"))[0] == 36 * 56 * 28 - 1


def test_boundary_shell_is_one_layer_box():
    cfg = scenes.shipped("small_block")
    _, bn, _ = scene.derive_sizes(cfg)
    bp = scene.init_boundary_positions(cfg, bn)
    assert len(np.unique(bp, axis=0)) == bn          # no duplicates
    on_wall = (np.isclose(bp[:, 0], 0) | np.isclose(bp[:, 0], 1.5) | np.isclose(bp[:, 1], 0) |
               np.isclose(bp[:, 1], 3.0) | np.isclose(bp[:, 2], 0) | np.isclose(bp[:, 2], 1.5))
    assert on_wall.all()


def test_voxelizer_on_a_synthetic_box(tmp_path):
    # 0.8 x 0.5 x 1.0 box (the shape of the reference's cube1.STL) written as an ASCII STL
    lo, hi = np.zeros(3), np.array([0.8, 0.5, 1.0])
    c = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    lines = ["solid box"]
    for q in quads:
        for tri in ((q[0], q[1], q[2]), (q[0], q[2], q[3])):
            lines.append("facet normal 0 0 0\nouter loop")
            lines += ["vertex %r %r %r" % tuple(float(t) for t in c[k]) for k in tri]
            lines.append("endloop\nendfacet")
    lines.append("endsolid box")
    p = tmp_path / "box.stl"
    p.write_text("\n".join(lines))
    v, f = scene.load_mesh(str(p))
    assert len(f) == 12 and np.allclose(v.min(0), lo) and np.allclose(v.max(0), hi)
    pts = scene.voxelize(v, f, 0.05, fill=True)
    assert pts.shape == (17 * 11 * 21, 3)            # SURVEY B-R1: 3 927 voxel points
    shell = scene.voxelize(v, f, 0.05, fill=False)
    assert len(shell) == 17 * 11 * 21 - 15 * 9 * 19


def test_read_config_roundtrip_and_exit_code(tmp_path):
    from cfd_taichi_b200 import utils
    import json
    p = tmp_path / "c.json"
    p.write_text(json.dumps(scenes.shipped("small_block")))
    assert utils.read_config(str(p))["solver"]["name"] == "dfsph"
    bad = tmp_path / "bad.json"
    bad.write_text("{ not json")
    r = subprocess.run([sys.executable, "-c",
                        "import sys; sys.path.insert(0, %r); from cfd_taichi_b200 import utils; utils.read_config(%r)"
                        % (ROOT, str(bad))], capture_output=True, text=True)
    assert r.returncode == 3                          # utils.py:8-11
    assert "Parsing config file error" in r.stdout


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sph_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libsph_b200.so does not export %s" % n
    assert sorted(p[0] for p in _lib.PROTOTYPES) == names, "ctypes prototypes out of sync with the header"
    assert L.sph_abi_version() == 1


def test_struct_layout_matches_header(built, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "sph_b200.h"\nint main(){printf("%zu %zu\\n", sizeof(SphConfig), sizeof(SphStats));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    a, b = (int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split())
    assert a == ctypes.sizeof(_lib.SphConfig) and b == ctypes.sizeof(_lib.SphStats)


def test_create_rejects_bad_arguments_without_gpu(built):
    L = _lib.load()
    h = ctypes.c_void_p()
    assert L.sph_create(None, 0, ctypes.byref(h)) == -1
    cfg = _lib.SphConfig()
    cfg.particle_radius = 0.025
    cfg.solver = 9
    for k in range(3):
        cfg.grid_num[k] = 4
    assert L.sph_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1
    assert b"unknown solver" in L.sph_last_error(None)


def test_no_cpu_fallback_in_product():
    # the product must never import the oracle, and must fail loudly without CUDA
    pkg = os.path.join(ROOT, "cfd_taichi_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in txt and "from oracle" not in txt, fn
    import torch
    if not torch.cuda.is_available():
        from cfd_taichi_b200.ParticleSystem import ParticleSystem
        with pytest.raises(_lib.SphError):
            ParticleSystem(scenes.shipped("small_block"))


def test_ply_and_obj_writers(tmp_path):
    # main.py:189-201: ASCII PLY with the vertex layout of ti.tools.PLYWriter, OBJ with 1-based faces
    from cfd_taichi_b200 import main as app
    pos = np.array([[0.1, 0.2, 0.3], [1.0, 2.0, 3.0]], dtype=np.float32)
    rgba = np.array([[0.0, 0.26, 0.68, 1.0]] * 2, dtype=np.float32)
    app.write_ply(str(tmp_path / "a.ply"), pos, rgba)
    lines = (tmp_path / "a.ply").read_text().splitlines()
    assert lines[:3] == ["ply", "format ascii 1.0", "element vertex 2"]
    assert [l.split()[-1] for l in lines[3:10]] == ["x", "y", "z", "red", "green", "blue", "alpha"] and lines[10] == "end_header"
    assert np.allclose(np.loadtxt(lines[11:]), np.concatenate([pos, rgba], axis=1), atol=1e-6)
    app.write_obj(str(tmp_path / "a.obj"), pos, np.array([[0, 1, 0]]))
    assert (tmp_path / "a.obj").read_text().splitlines() == ["v 0.100000 0.200000 0.300000", "v 1.000000 2.000000 3.000000", "f 1 2 1"]


def test_selfcheck_relinf_and_error_bits():
    """The comparison helper of the parity harness and the decoding of the device error flags (no GPU needed)."""
    import numpy as np
    import torch
    from cfd_taichi_b200 import _lib, selfcheck
    assert selfcheck.relinf([], []) == 0.0
    assert selfcheck.relinf([1.0, np.nan, 2.0], [1.0, np.nan, 2.0]) == 0.0            # equal garbage is no difference
    assert abs(selfcheck.relinf([1.0, 2.00002], [1.0, 2.0]) - 1e-5) < 1e-9
    assert selfcheck.relinf([np.nan], [1.0]) == float("inf")                          # a NaN where the reference has a number
    assert selfcheck.relinf(torch.tensor([1.0, 2.0]), np.array([1.0, 2.5])) == 0.2    # tensors and arrays mix
    assert selfcheck.relinf([1.0], [1.1], scale=10.0) == pytest.approx(0.01)
    err = {"a": {"x": 1e-7, "~info": {"x (vs current max)": 3e-3}}, "b": {"y": 2e-6}}
    assert selfcheck.worst(err) == (2e-6, "b / y")                                    # informational entries do not count
    assert _lib.decode_error_flags(0) == []
    msgs = _lib.decode_error_flags(2 | 32 | 64)
    assert len(msgs) == 3 and "max_neighbors" in msgs[0] and "peer" in msgs[1] and "bounds" in msgs[2]


def test_reference_golden_script_names_existing_scene_files():
    """tests/golden/make_reference_golden.py cannot run here (no taichi); at least its scene list must be real."""
    import importlib.util
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present")
    spec = importlib.util.spec_from_file_location("mrg", os.path.join(ROOT, "tests", "golden", "make_reference_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name, rel, solver in mod.CASES:
        assert os.path.exists(os.path.join(ref, rel)), rel
        assert solver in (None, "wcsph", "pcisph", "iisph", "dfsph")
    for solver, fields in mod.SOLVER_FIELDS.items():
        src = open(os.path.join(ref, solver + "_solver.py")).read() + open(os.path.join(ref, "solver_base.py")).read()
        for f in fields:
            assert "self.%s" % f in src, "%s_solver has no field %s" % (solver, f)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The bench contract, on the arm that runs without a GPU: stdout is ONE JSON line (anything a library writes to file
    descriptor 1 goes to stderr), carrying the base keys plus `impl`, `cpu_baseline` and a zero-copy `e2e`."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-n-side", "12"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "dfsph_particle_steps_per_sec" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # rank != 0 under torchrun: exits 0 without work and without output
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, cwd=root, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout == ""
