"""torchrun worker of tests/test_gpu_multigpu.py: one rank per GPU, x-slab run of a small jittered scene with
large x-velocities, gathered to rank 0 and compared with the single-domain run by global particle id
(cfd_taichi_b200/selfcheck.py:slab_vs_single -- the same check bench.py --gpus N reports in its `parity` block).

    torchrun ... mg_worker.py [steps] [strict|fast] [solver] [empty-edge]
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cfd_taichi_b200 import selfcheck  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    strict = (sys.argv[2] if len(sys.argv) > 2 else "strict") == "strict"
    solver = sys.argv[3] if len(sys.argv) > 3 else "dfsph"
    variant = sys.argv[4] if len(sys.argv) > 4 else ""
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cuts = None
    if variant == "empty-edge":
        # asymmetric halo counts.  The small block occupies the x-cell columns 3 .. 10 of 16; with the cut
        # after column 9 the last rank starts EMPTY while its neighbour's edge column is full: that neighbour sends,
        # receives nothing, and must still wait for the empty rank on every exchange.
        cuts = [0] + [3 + (8 * k) // (world - 1) for k in range(1, world - 1)] + [10, 16]
    res = selfcheck.slab_vs_single(solver=solver, steps=steps, strict=strict, cuts=cuts)
    ok = True
    if rank == 0:
        print("MGRESULT perm_ok=%s iters_ok=%s exact=%s dpos=%.3e dvel=%.3e migrated=%d two_cuts=%d owned_hist=%s flags=%s" % (
            res["every_particle_owned_once"], res["iters_ok"], res["slab_vs_single_bit_exact"], res["max_abs_dpos"],
            res["max_abs_dvel"], res["migrated_particles"], res["crossed_two_cuts"], res["owned_ghosts_last_step"],
            res["error_flags"]), flush=True)
        ok = res["ok"]
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
