#!/usr/bin/env python
"""bench.py -- headline benchmark: DFSPH particle-steps/s on the synthetic breaking dam.

  python bench.py --gpus N --steps K --warmup W           # this repository's CUDA path
  python bench.py --impl reference --gpus N --steps K ... # the restated reference on host cores

A "step" is one dfsph_solver.step() (grid build + neighbour lists + divergence-free solve +
non-pressure forces + constant-density solve + advection) over the whole particle block.
N = 1: BASELINE.json configs[1], 1 M particles (100^3) in a 15 x 8 x 5.2 box.
N > 1: configs[4], the dam is slab-decomposed along x, --n-side^3 particles per GPU (weak scaling; default
1 M per GPU, --n-side 200 = 8 M per GPU).

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier + synchronize on both
sides, max over ranks.  Three passes over the same K steps from one saved post-warm-up state: un-instrumented
(value), with an event pair around every launch (kernel_ms, roofline), end to end with host buffers (e2e).
Per-step working set (neighbour lists ~140 MB, gradient cache ~540 MB, 11 float4 arrays) exceeds the 126 MB
L2, so no explicit flush is needed between steps.

Besides the contract keys the line carries
  parity  N = 1: strict kernels bit-exact against the oracle and the fast kernels (the ones timed) sweep by sweep
          within 1e-5 of them, on the SAME 100^3 block, from the oracle's state after its warm-up steps;
          N > 1: the slab-decomposed run against the single-domain run, bit for bit by global particle id.
  step    whole-step algorithmic GB/s from SURVEY 8(d)'s B_step(D, C), neighbour_search_ms (grid build, lists)
  also    the other BASELINE configs timed in the same process: DFSPH 8 M (north_star's target size), WCSPH 30 k,
          PCISPH / IISPH 4 M (N = 1); DFSPH 8 M per GPU (N > 1, configs[4])
  comm    N > 1: ms per step in the halo-exchange kernels and in migration + ghost-particle exchange
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "dfsph_particle_steps_per_sec"
UNIT = "particle-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-side", type=int, default=100, help="particles per edge of the per-GPU block")
    ap.add_argument("--strict", action="store_true", help="strict-fp32 kernels (bit-exact vs the oracle)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline + parity leg (N = 1)")
    ap.add_argument("--cpu-n-side", type=int, default=0, help="edge of the CPU arm's block (default: --n-side, the GPU arm's block)")
    ap.add_argument("--no-also", action="store_true", help="skip the other BASELINE configs")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region: NVML in-process every 5 ms
    (nvidia-ml-py), falling back to the nvidia-smi query of the profiling recipe (one sample per ~50 ms)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []        # (sm_mhz, [bool x 4])
        self.sm_max = None
        self.source = None
        self.stop_flag = False

    def _run_nvml(self):
        import pynvml as N
        N.nvmlInit()
        # CUDA_VISIBLE_DEVICES may renumber the devices: address the GPU by its PCI bus id
        import torch
        bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
        hnd = None
        if bus is not None:
            for k in range(N.nvmlDeviceGetCount()):
                h = N.nvmlDeviceGetHandleByIndex(k)
                if int(N.nvmlDeviceGetPciInfo(h).bus) == int(bus):
                    hnd = h
                    break
        if hnd is None:
            hnd = N.nvmlDeviceGetHandleByIndex(self.index)
        self.sm_max = float(N.nvmlDeviceGetMaxClockInfo(hnd, N.NVML_CLOCK_SM))
        bits = [N.nvmlClocksEventReasonHwSlowdown, N.nvmlClocksEventReasonHwThermalSlowdown,
                N.nvmlClocksEventReasonSwThermalSlowdown, N.nvmlClocksEventReasonSwPowerCap]
        self.source = "nvml"
        while not self.stop_flag:
            r = N.nvmlDeviceGetCurrentClocksEventReasons(hnd)
            self.samples.append((float(N.nvmlDeviceGetClockInfo(hnd, N.NVML_CLOCK_SM)), [bool(r & b) for b in bits]))
            time.sleep(0.005)

    def _run_smi(self):
        self.source = "nvidia-smi"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.sm_max = float(parts[1])
                    self.samples.append((float(parts[0]), [p.lower().startswith("active") for p in parts[2:6]]))
            except Exception:
                pass
            time.sleep(0.05)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for k, n in enumerate(self.NAMES) if any(s[1][k] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(self.samples), "source": self.source}


def workload_name(n_side, n_gpus):
    n = n_side ** 3
    return "dfsph breaking dam, %d^3 = %d particles per GPU, r=0.025, Akinci boundary" % (n_side, n)


# ---------------------------------------------------------------------------------------------
# reference arm: the restated reference (oracle port, OpenMP) on the box's host cores
# ---------------------------------------------------------------------------------------------
def host_threads():
    return len(os.sched_getaffinity(0))


def run_reference(args):
    """`--impl reference`: the restated reference (oracle port, OpenMP over particles) on ALL host cores, on the GPU
    arm's own block (100^3 per GPU; for N > 1 one GPU's block -- CPU throughput does not depend on the block count)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O          # bench.py's cpu_baseline / reference legs may use oracle/
    from cfd_taichi_b200 import scenes
    n_side = args.cpu_n_side or args.n_side
    threads = host_threads()
    cfg = scenes.breaking_dam(n_side)
    o = O.Oracle(cfg, solver="dfsph", threads=threads)
    n = int(o.scalar("particle_num"))
    # bounded: ~4 s per 1 M-particle step on 16 cores; K + W capped so that the arm ends within a few minutes
    budget = max(3, int(30 * 1.0e6 / max(n, 1)))
    warm = min(args.warmup, max(1, budget // 5))
    steps = max(1, min(args.steps, budget - warm))
    for _ in range(warm):
        o.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step()
    dt = time.perf_counter() - t0
    info = {"divergence": int(o.scalar("df_div_iters")), "density": int(o.scalar("df_den_iters")), "of": "last timed step"}
    o.close()
    val, ms = n * steps / dt, dt / steps * 1e3
    sample = "%d^3 = %d-particle block (the GPU arm's per-GPU block), %d timed + %d warm-up dfsph steps, %d OpenMP threads" % (
        n_side, n, steps, warm, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.n_side, args.gpus), "sample": sample, "iterations": info,
                   "same_block_as_gpu_arm": n_side == args.n_side, "host_cores": threads},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
# algorithmic bytes per particle: per launch of one kernel class, and per step
ALG_BYTES = {"grid": 84, "lists": 20, "df_warm_start": 48, "df_drho": 28, "df_div_iter": 56, "df_ext_force": 40,
             "df_rho_adv": 32, "df_vel_adv_iter": 48, "df_position": 48}
STEP_BYTES = {"dfsph": lambda st: 268 + 84 * st.div_iters + 80 * st.den_iters, "wcsph": lambda st: 200,
              "pcisph": lambda st: 300 + 128 * st.pc_iters, "iisph": lambda st: 304 + 92 * st.ii_iters}
STEP_FORMULA = {"dfsph": "268 + 84 D + 80 C", "wcsph": "200", "pcisph": "300 + 128 I", "iisph": "304 + 92 I"}


_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def quiet(f, *a, **k):
    stdout = sys.stdout
    sys.stdout = open(os.devnull, "w")   # constructor prints (reference parity) must not pollute the JSON line
    try:
        return f(*a, **k)
    finally:
        sys.stdout.close()
        sys.stdout = stdout


def make_solver(cfg, solver, strict=False, slab=None):
    import importlib
    from cfd_taichi_b200.ParticleSystem import ParticleSystem
    cls = getattr(importlib.import_module("cfd_taichi_b200.%s_solver" % solver), "%s_solver" % solver)
    ps = quiet(ParticleSystem, cfg, strict=strict, solver_name=solver, slab=slab)
    return ps, quiet(cls, ps, cfg)


def iterations_of(solver, st):
    return {"dfsph": {"divergence": st.div_iters, "density": st.den_iters}, "pcisph": {"pressure": st.pc_iters},
            "iisph": {"pressure": st.ii_iters}, "wcsph": {}}[solver]


def traffic_per_particle(kernel):
    """DRAM bytes per particle of one launch, from the committed `ncu --set full` capture (profiles/traffic.json
    names the capture and the particle count it was taken at); scaled by the particles a launch processes here."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None, None
    with open(tpath) as f:
        t = json.load(f)
    rec = t.get("kernels", {}).get(kernel)
    if not rec:
        return None, None
    return float(rec["dram_bytes_per_launch"]) / float(rec["particles"]), rec.get("source", t.get("source"))


def measure_also(name, cfg, solver, warm, steps, barrier, world, rank, peak):
    """One of the other BASELINE configs, timed in this process with the same rules (CUDA events, barrier on both
    sides, max over ranks, working set > L2).  Returns the record for the line's `also` block (rank 0)."""
    import torch
    import torch.distributed as dist
    from cfd_taichi_b200 import _lib
    ps, sol = make_solver(cfg, solver, slab=(rank, world) if world > 1 else None)
    for _ in range(warm):
        sol.step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sol.step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    st = ps.read_stats()
    # per-kernel times of the same number of following steps
    L, h = ps._lib, ps._h
    _lib.check(L.sph_profile_begin(h), h)
    for _ in range(steps):
        sol.step()
    nk = len(_lib.KERNEL_CLASSES)
    ms_by, cnt_by = (ctypes.c_float * nk)(), (ctypes.c_int32 * nk)()
    _lib.check(L.sph_profile_end(h, ms_by, cnt_by, nk), h)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    n_total = ps.particle_num
    flags = st.error_flags
    ps.close()
    del ps, sol
    torch.cuda.empty_cache()
    bytes_step = STEP_BYTES[solver](st)
    gbps = bytes_step * (n_total / world) / (ms * 1e-3) / 1e9
    return {"workload": name, "solver": solver, "particles_total": n_total, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms, "particle_steps_per_sec": n_total / (ms * 1e-3), "iterations": iterations_of(solver, st),
            "error_flags": flags,
            "step_algorithmic_bytes_per_particle": bytes_step, "formula": STEP_FORMULA[solver],
            "achieved_GBps": gbps, "frac_of_hbm_peak": gbps / peak,
            "kernel_ms_per_step": {_lib.KERNEL_CLASSES[k]: round(float(ms_by[k]) / steps, 4) for k in range(nk) if cnt_by[k] > 0}}


def cpu_baseline_and_parity(args):
    """rank 0, N = 1.  The oracle port on all host cores on the GPU arm's own block: 2 warm-up + 3 timed steps
    (~20 s at 10^6 particles on 16 cores).  The state after the warm-up steps also feeds the `parity` block: the
    strict kernels walk the next step sweep by sweep (bit-exact against the oracle's next step), the fast kernels
    -- the ones this bench times -- are fed the strict inputs of every sweep and compared at 1e-5."""
    import numpy as np
    import torch
    from oracle import oracle as O          # bench.py's cpu_baseline / reference legs may use oracle/
    from cfd_taichi_b200 import scenes, selfcheck
    n_side = args.cpu_n_side or args.n_side
    threads = host_threads()
    cfg = scenes.breaking_dam(n_side)
    o = O.Oracle(cfg, solver="dfsph", threads=threads)
    n = int(o.scalar("particle_num"))
    warm, timed = 2, 3
    for _ in range(warm):
        o.step()
    parity = None
    if not args.no_parity:
        state = (o.field("pos").copy(), o.field("vel").copy(), o.field("warm_start_k").copy(), float(o.scalar("delta_time")))
        ps_s, sol_s = make_solver(cfg, "dfsph", strict=True)
        ps_f, sol_f = make_solver(cfg, "dfsph", strict=False)

        def load(ps, sol):
            ps.fluid_particles.pos.from_numpy(state[0])
            ps.fluid_particles.vel.from_numpy(state[1])
            ps._vel4[:n, 3] = torch.from_numpy(state[2]).to(ps._device)
            sol.delta_time[None] = state[3]

        load(ps_s, sol_s)
        selfcheck.copy_caller_state(ps_f, sol_f, ps_s, sol_s)
        err, info = selfcheck.sweeps("dfsph", ps_s, sol_s, ps_f, sol_f)
        strict_out = (ps_s.fluid_particles.pos.to_numpy(), ps_s.fluid_particles.vel.to_numpy(), sol_s.rho.to_numpy(),
                      ps_s.neighbour_counts().cpu().numpy())
        load(ps_f, sol_f)
        sol_f.step()                                   # the fused fast step, exactly what the timed region runs
        fast_out = (ps_f.fluid_particles.pos.to_numpy(), ps_f.fluid_particles.vel.to_numpy(), sol_f.rho.to_numpy())
        fast_iters = selfcheck.iters_of("dfsph", ps_f.read_stats())
        load(ps_s, sol_s)
        selfcheck.perturb_velocities_one_ulp(ps_s, seed=1)
        sol_s.step()                                   # the reference's own arithmetic on inputs moved by one ulp
        ulp_out = (ps_s.fluid_particles.pos.to_numpy(), ps_s.fluid_particles.vel.to_numpy())
        ps_s.close(); ps_f.close()
        del ps_s, ps_f, sol_s, sol_f
        torch.cuda.empty_cache()
    t0 = time.perf_counter()
    o.step()
    dt = time.perf_counter() - t0
    o_iters = (int(o.scalar("df_div_iters")), int(o.scalar("df_den_iters")))
    if not args.no_parity:
        ref = (o.field("pos"), o.field("vel"), o.field("rho"), o.field("nbr_count"))
        exact = {k: bool(np.array_equal(a, b)) for k, a, b in zip(("pos", "vel", "rho", "neighbour_count"), strict_out, ref)}
        w, where = selfcheck.worst(err)
        rel = selfcheck.relinf
        parity = {
            "block": "%d^3 = %d particles (the timed block), from the oracle's state after %d steps" % (n_side, n, warm),
            "oracle": "C restatement; reproduces, bit for bit, the reference's own solver sources executed under a stand-in for "
                      "the Taichi front end (tests/golden/refshim_*.npz, tests/test_reference_shim.py); unpinned for Taichi's "
                      "own code generation",
            "strict_vs_oracle": {"bit_exact": exact, "iterations": {"oracle": list(o_iters), "strict": list(info["iters"]["strict"])}},
            "fast_vs_strict_every_sweep": {"max_rel_err": w, "at": where, "tolerance": 1e-5, "sweeps": len(err),
                                           "loop_decisions_identical": info["loop_flags_equal"],
                                           "neighbour_counts_identical": info["neighbour_counts_equal"],
                                           "iterations": list(info["iters"]["fast"])},
            "fast_fused_substep_vs_oracle": {"rho": rel(fast_out[2], ref[2]), "pos": rel(fast_out[0], ref[0]),
                                             "vel": rel(fast_out[1], ref[1]), "iterations": list(fast_iters)},
            "strict_with_one_ulp_inputs_vs_oracle": {"pos": rel(ulp_out[0], ref[0]), "vel": rel(ulp_out[1], ref[1]),
                                                     "note": "the conditioning of the reference's own solver loops: what one "
                                                             "ulp on the input velocities does to one substep in strict arithmetic"},
        }
        parity["ok"] = bool(all(exact.values()) and tuple(info["iters"]["strict"]) == o_iters and w <= 1e-5
                            and info["loop_flags_equal"] and info["neighbour_counts_equal"]
                            and tuple(info["iters"]["fast"]) == o_iters and tuple(fast_iters) == o_iters
                            and parity["fast_fused_substep_vs_oracle"]["rho"] <= 1e-5
                            and parity["fast_fused_substep_vs_oracle"]["pos"] <= 1e-5)
    t0 = time.perf_counter()
    for _ in range(timed - 1):
        o.step()
    dt += time.perf_counter() - t0
    last = (int(o.scalar("df_div_iters")), int(o.scalar("df_den_iters")))
    o.close()
    base = {"value": n * timed / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d^3 = %d-particle block (%s), %d timed dfsph steps after %d warm-up steps, %d OpenMP threads "
                      "(div %d / den %d iterations in the last step)" % (
                          n_side, n, "the GPU arm's block" if n_side == args.n_side else "smaller than the GPU arm's block",
                          timed, warm, threads, last[0], last[1]),
            "same_block_as_gpu_arm": n_side == args.n_side, "ms_per_step": dt / timed * 1e3}
    return base, parity


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cfd_taichi_b200 import _lib, scenes, selfcheck

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path")
    torch.cuda.set_device(local_rank)
    def bind_near_gpu():
        # bind this thread to the CPUs next to its GPU, so that the pinned e2e buffers (first touch) and the copy engine's
        # host side sit on the GPU's NUMA node (a no-op on this pool's 16-CPU single-node boxes)
        before = os.sched_getaffinity(0)
        if os.environ.get("SPH_BENCH_NO_BIND"):      # A/B knob
            return before
        try:
            import pynvml as N
            N.nvmlInit()
            bus = int(torch.cuda.get_device_properties(local_rank).pci_bus_id)
            for k in range(N.nvmlDeviceGetCount()):
                hnd = N.nvmlDeviceGetHandleByIndex(k)
                if int(N.nvmlDeviceGetPciInfo(hnd).bus) == bus:
                    N.nvmlDeviceSetCpuAffinity(hnd)
                    break
        except Exception:
            pass
        return before

    if world > 1:
        bind_near_gpu()      # one process per GPU: for the whole run, instead of all ranks sharing one node
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1, before anything is timed: the slab-decomposed run against the single-domain run, bit for bit
    slab_parity = None
    if world > 1 and not args.no_parity:
        try:   # a failed check must be reported in the line, not cost the measurement
            slab_parity = selfcheck.slab_vs_single(solver="dfsph", steps=25, strict=True)
            fast_parity = selfcheck.slab_vs_single(solver="dfsph", steps=25, strict=False)   # the kernels this bench times
            if slab_parity is not None:
                slab_parity["fast_kernels"] = {k: fast_parity[k] for k in ("slab_vs_single_bit_exact", "iters_ok", "migrated_particles",
                                                                           "max_abs_dpos", "max_abs_dvel", "ok")}
                slab_parity["ok"] = bool(slab_parity["ok"] and fast_parity["ok"])
        except Exception as e:
            slab_parity = {"ok": False, "error": "%s: %s" % (type(e).__name__, e)} if rank == 0 else None

    # N > 1: the dam is `world` blocks long and slab-decomposed along x (weak scaling, configs[4])
    cfg = scenes.breaking_dam(args.n_side, gpus_x=world)
    ps, sol = make_solver(cfg, "dfsph", strict=args.strict, slab=(rank, world) if world > 1 else None)
    n_total = ps.particle_num
    L, h = ps._lib, ps._h

    for _ in range(max(args.warmup, 3)):
        sol.step()
    barrier()
    # The dam evolves (the density solve needs more iterations as it collapses), so every pass below --
    # timed, instrumented, end-to-end -- restarts from this same post-warm-up state and covers the same K steps.
    saved = (ps._pos4.clone(), ps._vel4.clone(), ps._gid.clone() if ps._gid is not None else None,
             ps.comm_info()["owned"] if world > 1 else None, sol.delta_time[None])

    def restore():
        ps._pos4.copy_(saved[0])
        ps._vel4.copy_(saved[1])
        if saved[2] is not None:
            ps._gid.copy_(saved[2])
            _lib.check(L.sph_set_counts(h, saved[3], 0), h)
        sol.delta_time[None] = saved[4]
        barrier()

    launches0 = ps.read_stats().kernel_launches

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        sol.step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    st_timed = ps.read_stats()
    launches = st_timed.kernel_launches - launches0
    restore()
    # Second pass over the same K steps with one CUDA-event pair around every launch (on the launching stream):
    # the per-kernel durations of the roofline.  The events themselves cost ~7 % of a step (two records per
    # launch, ~140 launches), so they are kept out of the pass `value` is taken from; the instrumented pass
    # reports its own ms per step next to it.
    _lib.check(L.sph_profile_begin(h), h)
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    evp0.record()
    for _ in range(args.steps):
        sol.step()
    evp1.record()
    barrier()
    ms_profiled = evp0.elapsed_time(evp1)
    nk = len(_lib.KERNEL_CLASSES)
    ms_by = (ctypes.c_float * nk)()
    cnt_by = (ctypes.c_int32 * nk)()
    _lib.check(L.sph_profile_end(h, ms_by, cnt_by, nk), h)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = n_total * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers in, host buffers out, through the C-ABI ------------------------------
    n = ps.particle_num if world == 1 else ps.comm_info()["owned"]
    ncap_e2e = n if world == 1 else ps._n_owned_cap
    # host state as the reference's callers hold it: pos / vel as N x 3 float32 (main.py:190), pinned
    affinity_before = bind_near_gpu() if world == 1 else None     # N = 1: for the e2e pass only (cpu_baseline uses every core)
    hpos = torch.empty((ncap_e2e, 3), dtype=torch.float32).pin_memory()
    hvel = torch.empty((ncap_e2e, 3), dtype=torch.float32).pin_memory()
    stream = ps._stream()
    e2e_steps = max(3, min(args.steps, 10))
    restore()
    _lib.check(L.sph_download_state_xyz(h, hpos.data_ptr(), hvel.data_ptr(), stream), h)
    for _ in range(2):       # untimed: first use of the upload path (copy stream, staging arrays, first touch of the pinned pages)
        _lib.check(L.sph_upload_state_xyz(h, hpos.data_ptr(), hvel.data_ptr(), stream), h)
        _lib.check(L.sph_step(h, 1, stream), h)
        _lib.check(L.sph_download_state_xyz(h, hpos.data_ptr(), hvel.data_ptr(), stream), h)
    restore()
    _lib.check(L.sph_download_state_xyz(h, hpos.data_ptr(), hvel.data_ptr(), stream), h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _lib.check(L.sph_upload_state_xyz(h, hpos.data_ptr(), hvel.data_ptr(), stream), h)
        _lib.check(L.sph_step(h, 1, stream), h)
        _lib.check(L.sph_download_state_xyz(h, hpos.data_ptr(), hvel.data_ptr(), stream), h)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / float(te.item())
    bytes_dir = 2 * n * 12
    if affinity_before is not None:
        os.sched_setaffinity(0, affinity_before)
    flags_main = ps.read_stats().error_flags
    if flags_main:
        raise SystemExit("bench.py: device error flags 0x%x: %s" % (flags_main, "; ".join(_lib.decode_error_flags(flags_main))))
    ps.close()
    del ps, sol, saved
    torch.cuda.empty_cache()

    if rank != 0 and os.environ.get("SPH_BENCH_ALL_RANKS"):
        sys.stderr.write("rank %d kernel_ms %s\n" % (rank, json.dumps({_lib.KERNEL_CLASSES[k]: round(float(ms_by[k]), 3)
                                                                       for k in range(nk) if cnt_by[k] > 0})))
    peak, peak_src = peaks()
    # ---- the other BASELINE configs, same process, same rules (every rank takes part when N > 1) -----------
    also = []
    if not args.no_also:
        if world == 1:
            todo = [("configs[4] block size: dfsph breaking dam, 200^3 = 8 M particles on one GPU (north_star's target size)",
                     scenes.breaking_dam(200), "dfsph", 3, 5),
                    ("configs[0]: config/breaking_dam_30k.json scene, solver wcsph", scenes.shipped("breaking_dam_30k", "wcsph"),
                     "wcsph", 20, 100),
                    ("configs[2]: pcisph breaking dam, 160^3 = 4.1 M particles", scenes.breaking_dam(160, "pcisph", 1.5e-4),
                     "pcisph", 400, 10),
                    ("configs[2]: iisph breaking dam, 160^3 = 4.1 M particles", scenes.breaking_dam(160, "iisph", 2.5e-4),
                     "iisph", 250, 10)]
        else:
            todo = [("configs[4]: dfsph breaking dam, 200^3 = 8 M particles per GPU, %d x-slabs" % world,
                     scenes.breaking_dam(200, gpus_x=world), "dfsph", 3, 5)]
        for name, acfg, solver, warm, steps in todo:
            try:
                rec = measure_also(name, acfg, solver, warm, steps, barrier, world, rank, peak)
            except Exception as e:     # an `also` line must never cost the headline
                rec = {"workload": name, "error": "%s: %s" % (type(e).__name__, e)}
            also.append(rec)

    if rank == 0:
        prof = {_lib.KERNEL_CLASSES[k]: {"ms": round(float(ms_by[k]), 4), "launches": int(cnt_by[k])}
                for k in range(nk) if cnt_by[k] > 0}
        # dominant kernel = largest share of the timed region
        compute = {k: v for k, v in prof.items() if not k.startswith("mg_")}
        dom = max(compute, key=lambda k: compute[k]["ms"])
        n_local = n_total / world
        avg_ms = prof[dom]["ms"] / prof[dom]["launches"]
        achieved = ALG_BYTES.get(dom, 0) * n_local / (avg_ms * 1e-3) / 1e9
        tpp, tsrc = traffic_per_particle(dom)
        traffic = tpp * n_local if tpp else None
        step_bytes = STEP_BYTES["dfsph"](st_timed)
        step_gbps = step_bytes * n_local / (ms_total / args.steps * 1e-3) / 1e9
        per_step = lambda k: round(prof[k]["ms"] / args.steps, 4) if k in prof else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.n_side, world), "particles_total": n_total,
                       "kernels": "strict-fp32" if args.strict else "fast-fp32",
                       "l2": "per-step working set ~900 MB > 126 MB L2, no flush",
                       "iterations": {"divergence": st_timed.div_iters, "density": st_timed.den_iters, "of": "last timed step"},
                       "parallelism": "1 GPU" if world == 1 else
                       "%d x-slabs; ghost values, loop partials, migration and ghost particles over CUDA-IPC peer windows "
                       "(NVLink)" % world},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_dir, "d2h_bytes_per_step": bytes_dir,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "traffic_gbps": (traffic / (avg_ms * 1e-3) / 1e9) if traffic else None,
                         "traffic_frac": (traffic / (avg_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "traffic_source": tsrc, "peak_source": peak_src,
                         "algorithmic_bytes_per_particle": ALG_BYTES.get(dom), "particles_per_launch": n_local,
                         "avg_launch_ms": avg_ms,
                         "timing": "CUDA events around every launch, second pass of the same %d steps "
                                   "(%.3f ms/step with the events in the stream)" % (args.steps, ms_profiled / args.steps)},
            "step": {"algorithmic_bytes_per_particle": step_bytes, "formula": STEP_FORMULA["dfsph"] + " (SURVEY 8(d))",
                     "achieved_GBps": step_gbps, "frac_of_hbm_peak": step_gbps / peak,
                     "neighbour_search_ms": {"grid_build": per_step("grid"), "neighbour_lists": per_step("lists")}},
            "kernel_ms": prof,
        }
        if world > 1:
            line["comm"] = {"exchange_ms_per_step": per_step("mg_exchange"), "begin_step_ms_per_step": per_step("mg_begin_step"),
                            "exchange_launches_per_step": prof.get("mg_exchange", {}).get("launches", 0) / args.steps,
                            "transport": "one kernel per exchange over CUDA-IPC peer windows (NVLink), migration and ghost "
                                         "particles included; no NCCL on the step path"}
            if slab_parity is not None:
                line["parity"] = slab_parity
        if world == 1 and not args.no_cpu_baseline:
            try:
                base, parity = cpu_baseline_and_parity(args)
            except Exception as e:   # the headline must survive a failure of the checker leg; the line says so
                base, parity = None, {"ok": False, "error": "%s: %s" % (type(e).__name__, e)}
            if base is not None:
                line["cpu_baseline"] = base
                line["vs_cpu_baseline"] = {"device_timed": value / base["value"], "e2e": e2e_value / base["value"],
                                           "host_cores": base["cores"],
                                           "note": "a reported baseline, not the target: it halves when the host has twice the cores"}
            if parity is not None:
                line["parity"] = parity
        if also:
            line["also"] = also
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global _JSON_OUT
    args = parse()
    # the contract is ONE JSON line on stdout: whatever a library writes to file descriptor 1 meanwhile (NCCL's version banner
    # at N > 1, constructor prints) is sent to stderr, and the line itself goes to the saved descriptor
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
