# coding=utf-8
"""main -- headless mirror of the reference entry point (main.py:13-211).

    python main.py --config <scene.json> [--steps N] [--output-dir DIR]

Same contract for everything on the hot path: the config schema (utils.read_config), the solver lookup
by name (module `<name>_solver`, class `<name>_solver`, main.py:65-68), one ParticleSystem, `iter_cnt`
solver steps then `iter_cnt` rigid steps per frame (main.py:166-171), simulated time advanced by
`iter_cnt * solver.delta_time[None]` (main.py:173), stop at t > 4.0 or 100 000 frames (main.py:98, 205),
PLY / OBJ export every 1 / output_fps of simulated time (main.py:189-200).  The GGUI window, camera and
keyboard handling (main.py:51-62, 108-161) need a display and are out of scope: this driver is headless
and defaults the camera keys the reference requires (SURVEY B-15).
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:          # lets importlib resolve "<name>_solver" as a top-level module name
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))

from cfd_taichi_b200 import utils  # noqa: E402
from cfd_taichi_b200.ParticleSystem import ParticleSystem  # noqa: E402
from cfd_taichi_b200.rigid_solver import rigid_solver  # noqa: E402


def write_ply(path, pos, rgba):
    """ASCII PLY with the vertex layout of ti.tools.PLYWriter (x y z red green blue alpha)."""
    n = pos.shape[0]
    with open(path, 'w') as f:
        f.write('ply\nformat ascii 1.0\nelement vertex %d\n' % n)
        for name in ('x', 'y', 'z', 'red', 'green', 'blue', 'alpha'):
            f.write('property float %s\n' % name)
        f.write('end_header\n')
        np.savetxt(f, np.concatenate([pos, rgba], axis=1), fmt='%.6f')


def write_obj(path, vertices, faces):
    with open(path, 'w') as f:
        for v in vertices:
            f.write('v %.6f %.6f %.6f\n' % tuple(v))
        if faces is not None:
            for t in faces:
                f.write('f %d %d %d\n' % (t[0] + 1, t[1] + 1, t[2] + 1))


def save_state(path, ps, solver, rs=None, t=0.0, frame_cnt=0, ply_cnt=0):
    """Restartable dump: everything a step depends on, plus the driver's own clock (simulated
    time, frame and output counters) so that a resumed run stops at the original end time and continues the
    output numbering."""
    import ctypes
    if ps._slab is not None:
        raise ValueError("save_state: a slab ParticleSystem holds one x-slab; dump from a single-domain run")
    n = ps.particle_num
    data = dict(pos4=ps._pos4[:n].cpu().numpy(), vel4=ps._vel4[:n].cpu().numpy(), delta_time=solver.delta_time[None],
                t=float(t), frame_cnt=int(frame_cnt), ply_cnt=int(ply_cnt))
    if ps.exist_rigid[None]:
        info = ps.rigid_state()
        data.update(rpos4=ps._rpos4.cpu().numpy(), rvel4=ps._rvel4.cpu().numpy(), rforce4=ps._rforce4.cpu().numpy(),
                    rverts4=ps._rverts4.cpu().numpy(),
                    rigid_info=np.frombuffer(ctypes.string_at(ctypes.addressof(info), ctypes.sizeof(info)), dtype=np.uint8).copy(),
                    centroid=np.array(list(info.centroid)), omega=np.array(list(info.omega)))
    np.savez(path, **data)


def load_state(path, ps, solver):
    """Restart from a save_state dump: positions, velocities (with the solver-persistent scalar in .w: DFSPH
    warm_start_k / IISPH p_past), the time step the next step starts from and, for scenes with a rigid body, its
    particles, mesh vertices and the device-side body state (sph_rigid_set_state)."""
    import ctypes
    import torch
    from cfd_taichi_b200 import _lib
    if ps._slab is not None:
        raise ValueError("load_state: a slab ParticleSystem holds one x-slab; restart from a single-domain run")
    path = path if path.endswith('.npz') else path + '.npz'
    clock = dict(t=0.0, frame_cnt=0, ply_cnt=0)
    with np.load(path) as d:
        for k in clock:
            if k in d:
                clock[k] = d[k].item()
        n = ps.particle_num
        if d['pos4'].shape[0] != n:
            raise ValueError("load_state: dump holds %d particles, the scene %d" % (d['pos4'].shape[0], n))
        ps._pos4[:n].copy_(torch.from_numpy(d['pos4']).to(ps._device))
        ps._vel4[:n].copy_(torch.from_numpy(d['vel4']).to(ps._device))
        solver.delta_time[None] = float(d['delta_time'])
        if ps.exist_rigid[None]:
            if 'rigid_info' not in d:
                raise ValueError("load_state: the dump holds no rigid-body state")
            for name, t in (('rpos4', ps._rpos4), ('rvel4', ps._rvel4), ('rforce4', ps._rforce4), ('rverts4', ps._rverts4)):
                t.copy_(torch.from_numpy(d[name]).to(ps._device))
            info = _lib.SphRigidInfo.from_buffer_copy(d['rigid_info'].tobytes())
            _lib.check(ps._lib.sph_rigid_set_state(ps._h, ctypes.byref(info)), ps._h)
    return clock


def check_error_flags(ps, frame_cnt):
    """The flags the library latches on the device mean the run has left the reference's physics (a truncated
    neighbour list, a clamped particle, the density-loop cap, a non-finite value, a dead peer): fail loudly."""
    from cfd_taichi_b200 import _lib
    flags = ps.read_stats().error_flags
    if flags:
        raise _lib.SphError("frame %d: device error flags 0x%x: %s" % (frame_cnt, flags, "; ".join(_lib.decode_error_flags(flags))))


def run(config, max_frames=None, output_dir='./output', quiet=False, resume=None, save=None, render_dir=None):
    scene_config = config.get('scene')
    solver_config = config.get('solver')
    print("Simulation Start!")
    start_time = time.time()
    ps = ParticleSystem(config)
    solver_name = solver_config.get('name')
    module = importlib.import_module('cfd_taichi_b200.' + solver_name + '_solver')     # main.py:65-68
    solver = getattr(module, solver_name + '_solver')(ps, config)
    rs = rigid_solver(ps, config) if config.get('solid', {}) else None                 # main.py:70-71
    clock = load_state(resume, ps, solver) if resume else dict(t=0.0, frame_cnt=0, ply_cnt=0)

    frame_cnt = int(clock['frame_cnt'])
    first_frame = frame_cnt
    iter_cnt = solver_config.get('iter_cnt')
    np_rgba = np.reshape(ps.rgba.to_numpy(), (ps.particle_num, 4))
    is_output_ply = scene_config.get('is_output_ply', False)
    output_fps = scene_config.get('output_fps', 60)
    frame_time = 1.0 / output_fps
    ply_cnt = int(clock['ply_cnt'])
    renderer_cnt = int(clock['ply_cnt'])
    t = float(clock['t'])
    if is_output_ply:
        os.makedirs(output_dir, exist_ok=True)
    renderer = None
    if render_dir:                                  # what window.show() / video_manager.write_frame show (main.py:175-188)
        from cfd_taichi_b200.render import Renderer
        os.makedirs(render_dir, exist_ok=True)
        renderer = Renderer(ps, config)
    while True:
        if frame_cnt > 100000 or (max_frames is not None and frame_cnt - first_frame >= max_frames):
            break
        for _ in range(iter_cnt):
            solver.step()
        for _ in range(iter_cnt):
            if rs and ps.active_rigid[None] == 1:
                rs.step()
        frame_cnt += 1
        check_error_flags(ps, frame_cnt)   # the same synchronising read the next line needs
        t += iter_cnt * solver.delta_time[None]
        if not quiet and frame_cnt % 50 == 0:
            print("time: {:.4f}  frame_cnt: {}  delta time: {:.5f}".format(t, frame_cnt, solver.delta_time[None]))
        if is_output_ply and (t / frame_time) > ply_cnt:
            np_pos = np.reshape(ps.fluid_particles.pos.to_numpy(), (ps.particle_num, 3))
            write_ply(os.path.join(output_dir, 'output_%06d.ply' % ply_cnt), np_pos, np_rgba)
            if ps.exist_rigid[None] == 1:
                ps.update_mesh_vextics()
                write_obj(os.path.join(output_dir, 'obj_%06d.obj' % ply_cnt), ps.mesh_vertices, ps._rigid_faces)
            ply_cnt += 1
        if renderer is not None and (t / frame_time) > renderer_cnt:
            renderer.save_png(os.path.join(render_dir, 'frame_%06d.png' % renderer_cnt))
            renderer_cnt += 1
        if t > 4.0:
            break
    if save:
        save_state(save, ps, solver, rs, t=t, frame_cnt=frame_cnt, ply_cnt=ply_cnt)
    print("Simulation time: {}".format(time.time() - start_time))
    return ps, solver, rs, t


def main(argv=None):
    parser = argparse.ArgumentParser(description='SPH on B200 (headless)')
    parser.add_argument('--config', help="Please input a scene config json file.", type=str, default='default.json')
    parser.add_argument('--steps', help="stop after this many frames (default: run to t > 4.0)", type=int, default=None)
    parser.add_argument('--output-dir', type=str, default='./output')
    parser.add_argument('--resume', help="restart from a state dump written with --save-state", type=str, default=None)
    parser.add_argument('--save-state', help="write a restartable .npz dump when the run ends", type=str, default=None)
    parser.add_argument('--render-dir', help="write a rendered frame (PNG, the scene file's camera) every 1 / output_fps of "
                        "simulated time: the headless stand-in for the reference's window", type=str, default=None)
    args = parser.parse_args(argv)
    config = utils.read_config(args.config)
    run(config, args.steps, args.output_dir, resume=args.resume, save=args.save_state, render_dir=args.render_dir)


if __name__ == "__main__":
    main()
