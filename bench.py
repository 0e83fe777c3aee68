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
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from cfd_taichi_b200 import _lib, scenes
    from cfd_taichi_b200.ParticleSystem import ParticleSystem
    from cfd_taichi_b200.dfsph_solver import dfsph_solver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # N > 1: the dam is `world` blocks long and slab-decomposed along x (weak scaling, configs[4])
    cfg = scenes.breaking_dam(args.n_side, gpus_x=world)
    devnull = open(os.devnull, "w")
    stdout = sys.stdout
    sys.stdout = devnull            # constructor prints (reference parity) must not pollute the JSON line
    try:
        ps = ParticleSystem(cfg, strict=args.strict, solver_name="dfsph", slab=(rank, world) if world > 1 else None)
        sol = dfsph_solver(ps, cfg)
    finally:
        sys.stdout = stdout
    n = ps.particle_num if world == 1 else ps.comm_info()["owned"]
    n_total = ps.particle_num
    L, h = ps._lib, ps._h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        sol.step()
    barrier()
    # The dam evolves (the density solve needs more iterations as it collapses), so every pass below --
    # timed, instrumented, end-to-end -- restarts from this same post-warm-up state and covers the same K steps.
    saved = (ps._pos4.clone(), ps._vel4.clone(), ps._gid.clone() if ps._gid is not None else None,
             ps.comm_info()["owned"] if world > 1 else None)

    def restore():
        ps._pos4.copy_(saved[0])
        ps._vel4.copy_(saved[1])
        if saved[2] is not None:
            ps._gid.copy_(saved[2])
            _lib.check(L.sph_set_counts(h, saved[3], 0), h)
        barrier()

    launches0 = ps.read_stats().kernel_launches

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        sol.step()
        # the DFSPH density loop already left the step's control block in pinned host memory
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    st_timed = ps.read_stats()
    launches = st_timed.kernel_launches - launches0
    restore()
    # Second pass over the next K steps of the same run with one CUDA-event pair around every launch (on the
    # launching stream): the per-kernel durations of the roofline.  The events themselves cost ~7 % of a
    # step (two records per launch, ~140 launches), so they are kept out of the pass `value` is taken from;
    # the instrumented pass reports its own ms per step next to it.
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
    st = ps.read_stats()

    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    total_particles = n_total
    value = total_particles * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers in, host buffers out, through the C-ABI ------------------------------
    n = ps.particle_num if world == 1 else ps.comm_info()["owned"]
    ncap_e2e = n if world == 1 else ps._n_owned_cap
    # host state as the reference's callers hold it: pos / vel as N x 3 float32 (main.py:190), pinned
    hpos = torch.empty((ncap_e2e, 3), dtype=torch.float32).pin_memory()
    hvel = torch.empty((ncap_e2e, 3), dtype=torch.float32).pin_memory()
    stream = ps._stream()
    _lib.check(L.sph_download_state_xyz(h, hpos.data_ptr(), hvel.data_ptr(), stream), h)
    e2e_steps = max(3, min(args.steps, 10))
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
    e2e_value = total_particles * e2e_steps / float(te.item())
    bytes_dir = 2 * n * 12

    if rank != 0 and os.environ.get("SPH_BENCH_ALL_RANKS"):
        sys.stderr.write("rank %d kernel_ms %s\n" % (rank, json.dumps({_lib.KERNEL_CLASSES[k]: round(float(ms_by[k]), 3)
                                                                       for k in range(nk) if cnt_by[k] > 0})))
    if rank == 0:
        prof = {_lib.KERNEL_CLASSES[k]: {"ms": round(float(ms_by[k]), 4), "launches": int(cnt_by[k])}
                for k in range(nk) if cnt_by[k] > 0}
        # dominant kernel = largest share of the timed region
        dom = max(prof, key=lambda k: prof[k]["ms"])
        # algorithmic bytes per particle per launch
        alg_bytes = {"grid": 84, "lists": 20, "df_warm_start": 48, "df_drho": 28, "df_div_iter": 56, "df_ext_force": 40,
                     "df_rho_adv": 32, "df_vel_adv_iter": 48, "df_position": 48}
        peak, peak_src = peaks()
        avg_ms = prof[dom]["ms"] / prof[dom]["launches"]
        achieved = alg_bytes.get(dom, 0) * (n_total / world) / (avg_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(dom)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.n_side, world), "particles_total": total_particles,
                       "kernels": "strict-fp32" if args.strict else "fast-fp32",
                       "l2": "per-step working set ~900 MB > 126 MB L2, no flush",
                       "iterations": {"divergence": st_timed.div_iters, "density": st_timed.den_iters, "of": "last timed step"},
                       "parallelism": "1 GPU" if world == 1 else
                       "%d x-slabs; per-sweep ghost values and loop partials over CUDA-IPC peer windows (NVLink), "
                       "NCCL for migration / ghost particles" % world},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_dir, "d2h_bytes_per_step": bytes_dir,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "traffic_gbps": (traffic / (avg_ms * 1e-3) / 1e9) if traffic else None,
                         "traffic_frac": (traffic / (avg_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_particle": alg_bytes.get(dom), "avg_launch_ms": avg_ms,
                         "timing": "CUDA events around every launch, second pass of the same %d steps "
                                   "(%.3f ms/step with the events in the stream)" % (args.steps, ms_profiled / args.steps)},
            "kernel_ms": prof,
        }
        if not args.no_cpu_baseline and world == 1:
            cn = min(args.cpu_n_side, args.n_side)
            cval, cms, threads, cnp, info = cpu_reference(cn, 2, 2)
            line["cpu_baseline"] = {
                "value": cval, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "%d^3 = %d-particle block of the same dam, 2 timed dfsph steps after 2 warm-up steps "
                          "(div %d / den %d iterations)" % (cn, cnp, info["div_iters"], info["den_iters"])}
        print(json.dumps(line), flush=True)
    ps.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
