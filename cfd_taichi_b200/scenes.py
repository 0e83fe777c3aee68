"""Scene table: the reference's shipped scenes and BASELINE.json's synthetic breaking-dam blocks,
built as config dicts in the reference JSON schema.  Users can pass
their own JSON files through utils.read_config; these builders exist so tests and bench.py do not
depend on files outside the repository.
"""
import copy


def make_scene(box_max, start_pos, water_size, solver, delta_time, boundary_handle=True, radius=0.025,
               gravity=9.8, solid=None, **solver_extra):
    cfg = {
        "scene": {"box_min": [0.0, 0.0, 0.0], "box_max": list(box_max), "particle_radius": radius,
                  "gravity": gravity, "is_output_gif": False, "is_output_ply": False, "is_simulate": True},
        "solver": dict({"name": solver, "delta_time": delta_time, "iter_cnt": 1,
                        "boundary_handle": boundary_handle}, **solver_extra),
        "fluid": {"start_pos": list(start_pos), "water_size": list(water_size)},
    }
    if solid is not None:
        cfg["solid"] = copy.deepcopy(solid)
    return cfg


_CUBE = {"mesh": "./obj/cube1.STL", "voxel_radius": 0.025, "scale": 1, "fill": True, "active": True}
_CUBE_LC = dict(_CUBE, mesh="./obj/cube1.stl")     # several shipped files spell the extension in lower case (SURVEY B-R5)

# name -> (box_max, start_pos, water_size, solver, dt, boundary_handle, solid block(s), extra scene keys)
# The values are those of the reference's default.json and config/*.json.  "solid1" is how
# default.json / breaking_dam_demo.json spell the key: the reference reads config.get('solid'), so those scenes run
# WITHOUT a rigid body -- kept as shipped.
_FPS = {"output_fps": 60}
_SHIPPED = {
    "default": ([7, 7, 2.5], [0.2, 0.1, 0.1], [2, 3.6, 2.3], "pcisph", 1e-3, False,
                {"solid1": dict(_CUBE_LC, rho_0=500, pos_offset=[4.7, 0.9, 0.7], attitude_offset=[0.0, 0.0, 90.0])},
                dict(_FPS, fs_couple=True)),
    "breaking_dam_30k": ([5.0, 3.0, 1.5], [0.1, 0.1, 0.1], [1.0, 2.8, 1.3], "iisph", 2.5e-4, True, None, {}),
    "breaking_dam_demo": ([10, 7, 3], [0.2, 0.1, 0.1], [2, 3.5, 2.8], "dfsph", 7e-4, False,
                          {"solid1": dict(_CUBE_LC, rho_0=500, pos_offset=[4.7, 0.9, 0.7], attitude_offset=[0.0, 0.0, 90.0])},
                          dict(_FPS, fs_couple=True)),
    "coupling_demo": ([5, 7, 2.5], [0.1, 0.1, 0.1], [1.5, 2.0, 2.3], "pcisph", 1e-4, True,
                      {"solid": dict(_CUBE_LC, rho_0=5000, pos_offset=[2.5, 0.9, 0.7], attitude_offset=[0.0, 0.0, 90.0])},
                      dict(_FPS, fs_couple=True)),
    "dam_flush_cube": ([5.0, 3.0, 1.5], [0.1, 0.1, 0.1], [1.8, 2.8, 1.4], "pcisph", 1e-4, True,
                       {"solid": dict(_CUBE, rho_0=2000, pos_offset=[3.0, 0.0, 0.2], attitude_offset=[0.0, 0.0, 0.0])}, {}),
    "experiment1_config": ([2.5, 2.4, 1.5], [0.1, 0.1, 0.1], [1.0, 2.0, 1.4], "iisph", 2.5e-4, True,
                           {"solid": dict(_CUBE_LC, rho_0=200, scale=0.6, pos_offset=[1.8, 0.0, 0.7],
                                          attitude_offset=[0.0, 0.0, 0.0])}, dict(_FPS, fs_couple=True)),
    "experiment2_config": ([2.5, 2.4, 1.5], [0.1, 0.1, 0.1], [1.0, 2.0, 1.4], "wcsph", 2.5e-4, True,
                           {"solid": dict(_CUBE_LC, rho_0=1000, scale=0.6, pos_offset=[1.7, 0.6, 0.7],
                                          attitude_offset=[0.0, 0.0, 90.0])}, dict(_FPS, fs_couple=True)),
    "small_block": ([1.5, 3.0, 1.5], [0.3, 0.5, 0.3], [0.7, 1.5, 0.7], None, None, True, None, {}),
}
# file name (without .json) -> (scene, solver): the five *_config_backup.json files are the small block per solver
FILES = {"default": ("default", None), "breaking_dam_30k": ("breaking_dam_30k", None),
         "breaking_dam_demo": ("breaking_dam_demo", None), "coupling_demo": ("coupling_demo", None),
         "dam_flush_cube": ("dam_flush_cube", None), "experiment1_config": ("experiment1_config", None),
         "experiment2_config": ("experiment2_config", None)}
FILES.update({s + "_config_backup": ("small_block", s) for s in ("dfsph", "iisph", "pbf", "pcisph", "wcsph")})

_SMALL_DT = {"dfsph": 1e-3, "iisph": 1e-3, "pbf": 2.5e-4, "pcisph": 1.5e-4, "wcsph": 5e-4}


def shipped(name, solver=None):
    """A shipped reference scene by name, optionally with solver.name overridden (BASELINE.json's
    configs override the solver named in breaking_dam_30k.json and dam_flush_cube.json)."""
    box, start, water, sol, dt, bh, solids, extra = _SHIPPED[name]
    if name == "small_block":          # the five *_config_backup.json scenes (N = 5879)
        sol = solver or "dfsph"
        dt = _SMALL_DT[sol]
    elif solver is not None:
        sol = solver
    extra = dict(extra)
    scene_extra = {k: extra.pop(k) for k in list(extra) if k == "output_fps"}
    cfg = make_scene(box, start, water, sol, dt, bh, **extra)
    cfg["scene"].update(scene_extra)
    for key, block in (solids or {}).items():
        cfg[key] = copy.deepcopy(block)
    return cfg


def breaking_dam(n_side, solver="dfsph", delta_time=1e-3, gpus_x=1):
    """Synthetic breaking-dam block of BASELINE.json: a cube of n_side^3 particles per GPU
    (100 -> 1 M, 160 -> 4.096 M, 200 -> 8 M) in a box 3x as long, at start_pos 0.1."""
    side = n_side * 0.05
    water = [side * gpus_x, side, side]
    box = [3.0 * side * gpus_x, 1.5 * side + (0.5 if n_side == 100 else 0.0), side + 0.2]
    if n_side == 100:
        box = [15 * gpus_x, 8, 5.2]
    elif n_side == 160:
        box = [24 * gpus_x, 12, 8.2]
    elif n_side == 200:
        box = [30 * gpus_x, 15, 10.2]
    return make_scene(box, [0.1, 0.1, 0.1], water, solver, delta_time, True)
