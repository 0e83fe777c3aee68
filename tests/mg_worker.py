"""torchrun worker of tests/test_gpu_multigpu.py: one rank per GPU, x-slab DFSPH, results gathered to
rank 0 and compared with the single-domain run by global particle id."""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cfd_taichi_b200 import scenes  # noqa: E402
from cfd_taichi_b200.ParticleSystem import ParticleSystem  # noqa: E402
import importlib  # noqa: E402


def random_state(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-0.008, 0.008, size=(n, 3)).astype(np.float32),
            (rng.normal(0, 1.0, size=(n, 3)) * np.array([3.0, 0.5, 0.5])).astype(np.float32))


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    strict = (sys.argv[2] if len(sys.argv) > 2 else "strict") == "strict"
    solver = sys.argv[3] if len(sys.argv) > 3 else "dfsph"
    solver_cls = getattr(importlib.import_module("cfd_taichi_b200.%s_solver" % solver), "%s_solver" % solver)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = scenes.shipped("small_block", solver)
    n_global = 5879
    jit, vel = random_state(n_global, 7)
    with contextlib.redirect_stdout(io.StringIO()):
        ps = ParticleSystem(cfg, strict=strict, solver_name=solver, slab=(rank, world))
        sol = solver_cls(ps, cfg)
    gid, pos, v4 = ps.owned_state()
    n0 = len(gid)
    ps._pos4[:n0, :3] += torch.from_numpy(jit[gid]).to(ps._device)
    ps._vel4[:n0, :3] = torch.from_numpy(vel[gid]).to(ps._device)
    migrated = 0
    hist = []
    for _ in range(steps):
        sol.step()
        info = ps.comm_info()
        hist.append((info["owned"], info["ghosts"]))
    st = sol.stats()
    gid, pos, v4 = ps.owned_state()
    rho = sol.rho.to_torch()[:len(gid)].cpu().numpy()
    out = dict(rank=rank, gid=gid, pos=pos, vel=v4, rho=rho, hist=hist, div=st.div_iters, den=st.den_iters, pc=st.pc_iters, ii=st.ii_iters,
               dt=st.delta_time, flags=st.error_flags)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(out, gathered, dst=0)
    ok = True
    if rank == 0:
        with contextlib.redirect_stdout(io.StringIO()):
            ps1 = ParticleSystem(cfg, strict=strict, solver_name=solver)
            sol1 = solver_cls(ps1, cfg)
        ps1._pos4[:n_global, :3] += torch.from_numpy(jit).to(ps1._device)
        ps1._vel4[:n_global, :3] = torch.from_numpy(vel).to(ps1._device)
        for _ in range(steps):
            sol1.step()
        st1 = sol1.stats()
        ref_pos, ref_vel = ps1._pos4[:n_global, :3].cpu().numpy(), ps1._vel4[:n_global].cpu().numpy()
        gids = np.concatenate([g["gid"] for g in gathered])
        pos = np.concatenate([g["pos"] for g in gathered])
        vel4 = np.concatenate([g["vel"] for g in gathered])
        perm_ok = np.array_equal(np.sort(gids), np.arange(n_global))
        order = np.argsort(gids)
        pos, vel4 = pos[order], vel4[order]
        moved = sum(abs(g["hist"][-1][0] - g["hist"][0][0]) for g in gathered)
        iters_ok = all((g["div"], g["den"], g["pc"], g["ii"]) == (st1.div_iters, st1.den_iters, st1.pc_iters, st1.ii_iters)
                       for g in gathered)
        dpos = float(np.abs(pos - ref_pos).max())
        dvel = float(np.abs(vel4 - ref_vel).max())
        exact = np.array_equal(pos, ref_pos) and np.array_equal(vel4, ref_vel)
        print("MGRESULT perm_ok=%s iters_ok=%s exact=%s dpos=%.3e dvel=%.3e owned_hist=%s flags=%s" % (
            perm_ok, iters_ok, exact, dpos, dvel, [g["hist"][-1] for g in gathered], [g["flags"] for g in gathered]),
            flush=True)
        tol_ok = exact if strict else (dpos <= 1e-3 and perm_ok)
        ok = perm_ok and iters_ok and tol_ok and all(g["flags"] == 0 for g in gathered)
        ps1.close()
    ps.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
