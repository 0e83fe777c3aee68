"""CPU tests of the multi-GPU host logic (slab planning, lattice partition, ghost-layer protocol).
The exchange protocol is exercised with world_size = 2 over gloo; the neighbour counts of every owned
particle computed from (owned + one ghost column) must equal the single-domain oracle's."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cfd_taichi_b200 import scene, scenes, slab
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_cuts_properties():
    rng = np.random.default_rng(0)
    for _ in range(200):
        gx = int(rng.integers(8, 400))
        hist = rng.integers(0, 1000, size=gx) * (rng.random(gx) < 0.5)
        for p in (1, 2, 4, 8):
            if gx < p:
                continue
            cuts = slab.plan_cuts(hist, p)
            assert cuts[0] == 0 and cuts[-1] == gx and len(cuts) == p + 1
            assert all(b > a for a, b in zip(cuts, cuts[1:]))          # every rank owns >= 1 column


def test_breaking_dam_is_balanced_not_equal_width():
    cfg = scenes.breaking_dam(100, gpus_x=4)
    hist, _ = slab.column_histogram(cfg)
    cuts = slab.plan_cuts(hist, 4)
    counts = [int(hist[a:b].sum()) for a, b in zip(cuts, cuts[1:])]
    assert sum(counts) == 4000000 and max(counts) - min(counts) <= int(hist.max())
    # the fluid occupies the first third of the box: equal-width slabs would leave 2 of 4 ranks empty
    assert cuts[1] < len(hist) // 4


@pytest.mark.parametrize("nranks", [2, 3])
def test_lattice_partition_is_exact(nranks):
    cfg = scenes.shipped("small_block", "dfsph")
    n = scene.derive_sizes(cfg)[0]
    full = scene.init_fluid_positions(cfg, n)
    hist, _ = slab.column_histogram(cfg)
    assert hist.sum() == n                                    # exact, including the missing last site (B-1)
    cuts = slab.plan_cuts(hist, nranks)
    seen = []
    for r in range(nranks):
        ids, _, _ = slab.owned_lattice_ids(cfg, cuts[r], cuts[r + 1])
        pos = scene.init_fluid_positions(cfg, n, ids)
        assert np.array_equal(pos, full[ids])
        col = np.floor(pos[:, 0] / np.float32(0.1)).astype(int)
        assert col.min() >= cuts[r] and col.max() < cuts[r + 1]
        owned0, owned_cap, ghost_cap = slab.capacities(cfg, cuts, r)
        assert owned_cap >= len(ids) and ghost_cap > 0
        seen.append(ids)
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(n))


def _neighbour_counts(centres, others, h):
    """fp32 restatement of the cull (PS:466): sqrt((dx*dx + dy*dy) + dz*dz) > h excludes."""
    out = np.zeros(len(centres), dtype=np.int32)
    for k, c in enumerate(centres):
        d = c[None, :] - others
        r = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        out[k] = int((r <= h).sum()) - 1                      # minus the particle itself
    return out


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = scenes.shipped("small_block", "dfsph")
    n = scene.derive_sizes(cfg)[0]
    rng = np.random.default_rng(3)
    jitter = rng.uniform(-0.01, 0.01, size=(n, 3)).astype(np.float32)
    hist, _ = slab.column_histogram(cfg)
    cuts = slab.plan_cuts(hist, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    ids, _, _ = slab.owned_lattice_ids(cfg, lo, hi)
    pos = scene.init_fluid_positions(cfg, n, ids) + jitter[ids]
    h = np.float32(0.1)
    # migration: jitter moved some particles across the cut -> hand them to the neighbour
    col = np.floor(pos[:, 0] / h).astype(np.int64)
    cap = 4096

    def exchange(send_left, send_right):
        """Fixed-capacity messages with the count in the header (the protocol of sph_multigpu.cu)."""
        got = []
        for peer, payload in ((rank - 1, send_left), (rank + 1, send_right)):
            if peer < 0 or peer >= world:
                got.append(np.zeros((0, 4), dtype=np.float32))
                continue
            msg = torch.zeros((cap + 1, 4), dtype=torch.float32)
            msg[0, 0] = len(payload)
            msg[1:1 + len(payload)] = torch.from_numpy(payload)
            rcv = torch.zeros((cap + 1, 4), dtype=torch.float32)
            ops = [dist.P2POp(dist.isend, msg, peer), dist.P2POp(dist.irecv, rcv, peer)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            got.append(rcv[1:1 + int(rcv[0, 0])].numpy())
        return got

    rec = np.concatenate([pos, ids[:, None].astype(np.float32)], axis=1)
    go_l = rec[(col < lo) & (rank > 0)]
    go_r = rec[(col >= hi) & (rank + 1 < world)]
    keep = rec[~(((col < lo) & (rank > 0)) | ((col >= hi) & (rank + 1 < world)))]
    from_l, from_r = exchange(go_l, go_r)
    owned = np.concatenate([keep, from_l, from_r])
    col = np.floor(owned[:, 0] / h).astype(np.int64)
    assert col.min() >= (lo if rank > 0 else 0) and (col.max() < hi or rank + 1 == world)
    # ghost layer: my first / last owned column goes to the left / right neighbour
    gl, gr = exchange(owned[col == lo], owned[col == hi - 1])
    allp = np.concatenate([owned, gl, gr])
    counts = _neighbour_counts(owned[:, :3], allp[:, :3], h)
    out = [None] * world if rank == 0 else None
    dist.gather_object((owned[:, 3].astype(np.int64), counts), out, dst=0)
    if rank == 0:
        o = O.Oracle(cfg, solver="dfsph", threads=2)
        p = o.field("pos")
        p += jitter
        o.phase("reset_grid_update_grid")
        o.phase("neighbour_counts")
        ref = o.field("nbr_count").copy()
        gids = np.concatenate([g for g, _ in out])
        cnts = np.concatenate([c for _, c in out])
        ok = np.array_equal(np.sort(gids), np.arange(n)) and np.array_equal(cnts[np.argsort(gids)], ref)
        ret.put(bool(ok))
        o.close()
    dist.barrier()
    dist.destroy_process_group()


def test_ghost_layer_protocol_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 29000 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert ret.get() is True


def test_global_argmax_equals_single_domain():
    # PCISPH picks its delta particle with the reference's racy arg-max (PS:409-422); over slabs the decision
    # is taken on the merged counts and must name the same GLOBAL particle, whoever owns it
    from cfd_taichi_b200 import slab
    rng = np.random.default_rng(11)
    for _ in range(20):
        n = int(rng.integers(1, 400))
        counts = rng.integers(0, 6, size=n)
        want = slab.racy_argmax(counts)
        perm = rng.permutation(n)
        cut = sorted(rng.integers(0, n + 1, size=2))
        ranks = [perm[:cut[0]], perm[cut[0]:cut[1]], perm[cut[1]:]]
        k, owner = slab.global_argmax(n, [(g, counts[g]) for g in ranks])
        assert k == want and (k == -1 or k in set(ranks[owner].tolist()))
    assert slab.racy_argmax([]) == -1 and slab.global_argmax(0, []) == (-1, -1)
    # one-thread semantics: atomic_max returns the old maximum, so only a REPEATED maximum registers
    assert slab.racy_argmax([3, 1, 3, 2, 3]) == 4 and slab.racy_argmax([1, 2, 3]) == -1
