"""Host-side planning of the x-slab decomposition: which x-cell columns each rank
owns, which lattice particles start on it, and how much room it needs.  Pure numpy (no GPU, no
collectives): the exchange itself lives in csrc/sph_multigpu.cu; torch.distributed only carries the
128-byte NCCL id and test gathers."""
import math

import numpy as np

from . import scene


def column_histogram(config):
    """Particles per x-cell column of the initial fluid lattice (PS:142-150, 490-494)."""
    sc, fl = config["scene"], config["fluid"]
    r = sc["particle_radius"]
    d, h = r * 2, np.float32(4 * r)
    x_num = int(round(fl["water_size"][0] / d))
    n_total, _, grid = scene.derive_sizes(config)
    xi = np.arange(x_num, dtype=np.int64)
    x = (xi.astype(np.float32) * np.float32(r)) * np.float32(2.0) + np.float32(fl["start_pos"][0])
    col = np.floor(x / h).astype(np.int64)
    # particles per lattice x index: the lattice is filled x-fastest and may end mid-row (SURVEY B-1)
    per_x = n_total // x_num + (xi < n_total % x_num).astype(np.int64)
    hist = np.zeros(grid[0], dtype=np.int64)
    np.add.at(hist, col, per_x)
    return hist, col


def plan_cuts(hist, nranks):
    """Column cuts c_0 = 0 < c_1 < ... < c_P = G_x with (nearly) equal particle counts per slab.
    A breaking dam starts in one corner, so equal-width slabs would leave most ranks empty."""
    gx = len(hist)
    total = int(hist.sum())
    cum = np.concatenate([[0], np.cumsum(hist)])
    cuts = [0]
    for p in range(1, nranks):
        target = total * p / nranks
        c = int(np.searchsorted(cum, target, side="left"))
        c = max(c, cuts[-1] + 1)                   # at least one column per rank
        c = min(c, gx - (nranks - p))
        cuts.append(c)
    cuts.append(gx)
    return cuts


def owned_lattice_ids(config, col_lo, col_hi):
    """Global lattice indices (gid = x + x_num * (z + z_num * y), the reference's particle index) of
    the particles that start in columns [col_lo, col_hi)."""
    sc, fl = config["scene"], config["fluid"]
    d = sc["particle_radius"] * 2
    n_total = scene.derive_sizes(config)[0]
    x_num = int(round(fl["water_size"][0] / d))
    z_num = int(round(fl["water_size"][2] / d))
    _, col = column_histogram(config)
    xs = np.nonzero((col >= col_lo) & (col < col_hi))[0].astype(np.int64)
    rows = np.arange((n_total + x_num - 1) // x_num, dtype=np.int64)     # (y, z) rows
    gid = (rows[:, None] * x_num + xs[None, :]).reshape(-1)
    gid = gid[gid < n_total]
    return np.sort(gid), x_num, z_num


def capacities(config, cuts, rank):
    """(owned capacity, ghost capacity) for one rank: room for the flow to pile up in its slab and for
    two full boundary columns per side."""
    hist, _ = column_histogram(config)
    lo, hi = cuts[rank], cuts[rank + 1]
    owned0 = int(hist[lo:hi].sum())
    col_max = int(hist.max()) if hist.size else 0
    owned_cap = int(math.ceil(owned0 * 1.5)) + 4 * col_max + 65536
    ghost_cap = 2 * (3 * col_max + 16384)
    return owned0, owned_cap, ghost_cap


def racy_argmax(counts):
    """get_max_neighbor_particle_index (PS:409-422) with ti.cpu / one thread: atomic_max returns the OLD
    maximum, so the winner is the LAST index whose count equals the running maximum of the counts before it."""
    counts = np.asarray(counts, dtype=np.int64)
    if counts.size == 0:
        return -1
    run = np.maximum.accumulate(np.concatenate([[-1], counts[:-1]]))
    hit = np.nonzero(run == counts)[0]
    return int(hit[-1]) if hit.size else -1


def global_argmax(total, parts):
    """The same decision for a slab-decomposed domain.  `parts` = one (gids, counts) pair per rank (owned
    particles only).  Returns (global index of the winner, rank that owns it); (-1, -1) for an empty domain."""
    glob = np.zeros(total, dtype=np.int64)
    owner = np.full(total, -1, dtype=np.int64)
    for r, (g, c) in enumerate(parts):
        g = np.asarray(g, dtype=np.int64)
        glob[g] = np.asarray(c, dtype=np.int64)
        owner[g] = r
    k = racy_argmax(glob)
    return (k, int(owner[k])) if k >= 0 else (-1, -1)
