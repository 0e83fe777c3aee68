"""GPU parity of the multi-GPU slab path (BASELINE.json configs[4]): 2 / 4 / 8 ranks vs the single-domain run,
compared by global particle id.  Needs that many GPUs; skipped otherwise (the driver's test box has one GPU:
bench.py --gpus N runs the same check -- selfcheck.slab_vs_single -- and reports it in its `parity` block)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, args, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "mg_worker.py")] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_two_slabs_match_single_domain(built, mode):
    r = _run(2, ["25", mode])
    line = [l for l in r.stdout.splitlines() if l.startswith("MGRESULT")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert line and "perm_ok=True" in line[0] and "iters_ok=True" in line[0]
    # bit-identical to the single-GPU run of the same arithmetic mode, migration included -- strict AND fast: the
    # sort tie-break is the global id, the sums run in the same order, and the control arithmetic (adaptive dt)
    # uses explicit IEEE intrinsics in whichever translation unit decides
    assert "exact=True" in line[0]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("solver", ["wcsph", "pcisph", "iisph", "pbf"])
def test_two_slabs_other_solvers_match_single_domain(built, solver):
    # WCSPH: one ghost-density exchange per step.  PCISPH: ghost density, then per pressure iteration the ghosts'
    # pressure and predicted position plus the residual all-reduce.  Bit-identical to the single-domain run.
    r = _run(2, ["25", "strict", solver])
    line = [l for l in r.stdout.splitlines() if l.startswith("MGRESULT")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert line and "perm_ok=True" in line[0] and "iters_ok=True" in line[0] and "exact=True" in line[0]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_slabs_with_rigid_body(built):
    # replicated rigid body, fluid->rigid forces as partial sums + one all-reduce per rigid step:
    # every rank integrates the identical body state; fluid and body agree with the single-domain run to tolerance
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tests", "mg_worker_rigid.py"), "40"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("MGRIGID")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert line and "perm_ok=True" in line[0] and "replicas_identical=True" in line[0]


def _need(n):
    return pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < n, reason="needs %d GPUs" % n)


@_need(2)
@pytest.mark.parametrize("solver", ["wcsph", "dfsph"])
def test_two_slabs_with_an_empty_edge_column(built, solver):
    """asymmetric halo counts.  The last rank starts empty, so its neighbour sends to it without ever
    receiving from it; every exchange must still handshake with it (WCSPH has no all-rank reduce at all), or the
    sender runs two epochs ahead and overwrites unread slots."""
    r = _run(2, ["40", "strict", solver, "empty-edge"])
    line = [l for l in r.stdout.splitlines() if l.startswith("MGRESULT")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert line and "perm_ok=True" in line[0] and "iters_ok=True" in line[0] and "exact=True" in line[0]


@pytest.mark.parametrize("nranks", [4, 8])
def test_many_slabs_match_single_domain(built, nranks):
    """4 and 8 slabs of the 14-column block (three to four / one to two columns per rank, i.e. ranks that send the
    same particle to both neighbours): hundreds of particles migrate.  (The solver damps the initial velocities
    within a few steps, so no particle gets across two cuts in 60 steps; the line reports the count.)"""
    if not torch.cuda.is_available() or torch.cuda.device_count() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    r = _run(nranks, ["60", "strict", "dfsph"])
    line = [l for l in r.stdout.splitlines() if l.startswith("MGRESULT")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert line and "perm_ok=True" in line[0] and "iters_ok=True" in line[0] and "exact=True" in line[0]
    migrated = int(line[0].split("migrated=")[1].split()[0])
    two_cuts = int(line[0].split("two_cuts=")[1].split()[0])
    assert migrated > 100 and two_cuts >= 0, line[0]


@_need(2)
@pytest.mark.parametrize("knob", ["SPH_MG_OVERLAP", "SPH_MG_EPILOGUE_PUSH", "SPH_MG_TRANSPORT"])
@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_two_slabs_opt_in_paths_stay_bit_identical(built, knob, mode):
    """The measured-and-not-adopted variants (exchange behind the interior launch, edge values
    pushed from the sweeps' epilogue) and the NCCL transport kept for A/B runs must stay correct while they exist."""
    r = _run(2, ["25", mode], env={knob: "nccl" if knob == "SPH_MG_TRANSPORT" else "1"})
    line = [l for l in r.stdout.splitlines() if l.startswith("MGRESULT")]
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert line and "perm_ok=True" in line[0] and "iters_ok=True" in line[0] and "exact=True" in line[0]
