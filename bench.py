#!/usr/bin/env python
"""Benchmark of the photon-packet loop (BASELINE.json metric: photon packets/sec on the
ref4.1 disk at 1/2/4/8 B200 vs host OpenMP).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm
    python bench.py --impl reference --gpus N --steps K ...   # host-core OpenMP arm

One "step" = one BLOCKING thermal-mode pass of mc_photon_loop (dust_transfer.f90:439) over one batch of
packets on the synthetic ref4.1-like model G1: 128 chunks x --n2 packets per GPU (weak scaling:
chunks are dealt round-robin to ranks, n2 is scaled by the number of ranks).  What a drop-in caller gets:
one call at a time, nothing overlapped.  Besides the headline the line carries
  sweep[]   the same call at 1.28e5 .. 1.28e8 packets (the reference arm runs the SAME budgets),
  strong    (N > 1) the fixed 1.28e8-packet job split over the N ranks.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "photon packets/sec (ref4.1 disk, thermal mc_photon_loop)"
UNIT = "packets/s"
WORKLOAD = "G1 ref4.1-like: cylindrical 100x70x1, 50 lambda, n_T=100, thermal step, tau_mid(0.81um)=1e5, dark zone tau>1500"
# SURVEY 8d algorithmic bytes: per cell-crossing step (2D cyl), per scattering, per absorption, per packet
B_STEP, B_SCA, B_ABS = 112.0, 72.0, 112.0
SWEEP_N2 = (1000, 10000, 100000, 1000000)          # x 128 chunks: 1.28e5 (the .para budget) .. 1.28e8 packets


def b_packet(n_cells):
    return np.ceil(np.log2(n_cells)) * 8.0 + 96.0


def _json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


# ref4.1.para: "T T" = separation of contributions, Stokes parameters (read_param.f90:163);
# lsepar_pola stays on during the thermal step, so scatterings update the Stokes vector.
FLAGS = dict(lsepar_pola=1, lsepar_contrib=1)


# the other BASELINE configs (bench lines on request: --workload g2|g4|g5; the driver's default line is G1)
WORKLOADS = {
    "g1": dict(name=WORKLOAD, b_step=112.0, flags={}, n2=1000000),
    "g2": dict(name="G2 Pascucci-like (Pascucci_3.0.para): cylindrical 100x70x1, 61 lambda, one 0.12 um grain, isotropic scattering, tau_V = 100, thermal step",
               b_step=112.0, flags=dict(lisotropic=1), n2=10000),
    "g3": dict(name="G3 ref4.1_multi-like (LTE part): cylindrical 100x70x1, two zones with different dust, every opacity / phase-function / emission table per cell (280 MB of emission CDFs in global memory), 50 lambda, thermal step, tau_mid(0.81um)=1e5",
               b_step=128.0, flags=dict(lsepar_pola=0, lsepar_contrib=0), n2=100000),
    "g4": dict(name="G4 ref4.1_3D-like: cylindrical 100x50x72 two-sided = 720 000 cells, 50 lambda, thermal step, tau_mid(0.81um)=1e3",
               b_step=120.0, flags=dict(lsepar_pola=0, lsepar_contrib=0), n2=100000),
    "g5": dict(name="G5 Voronoi mesh of a synthetic 1M-particle SPH disk (997 016 cells, 15.5 neighbours per cell), 50 lambda, thermal step, tau_mid=1e3",
               b_step=332.0, flags=dict(lsepar_pola=0, lsepar_contrib=0), n2=20000),
}


_LANE = "mc_photon_loop_kernel<%s,%d,BANK,0> (packet per lane; calls above 1.2e6 packets end with a launch of mc_warp_engine_kernel, packet per warp, on the parked stragglers; smaller calls run on mc_warp_engine_kernel alone)"
KERNELS = {"g1": _LANE % ("GeomCyl<0,1>", 1), "g2": _LANE % ("GeomCyl<0,1>", 1), "g3": _LANE % ("GeomCyl<0,0>", 0),
           "g4": _LANE % ("GeomCyl<1,1>", 1), "g5": _LANE % ("GeomVor", 0)}
_RESIDENT = ("tables and tallies are L2 / shared-memory resident (DRAM traffic ~ 1e-4 of the algorithmic bytes): the ceiling is the L2 reduction "
             "rate and the issue rate of a divergent fp64 code, not HBM; hbm_frac is the formal algorithmic-bytes figure")
_STREAMED = ("per-cell data (%s) exceed shared memory and are read through L2 (126 MB) / HBM with data-dependent addresses: "
             "roofline against the measured copy bandwidth with SURVEY 8d's algorithmic bytes per step")
NOTES = {"g1": _RESIDENT, "g2": _RESIDENT,
         "g3": "per-cell opacities / phase functions (7000 cells) are L2-resident, the 280 MB of per-cell emission CDFs are touched 3.8 times per packet "
               "(measured DRAM traffic 3e-4 of the algorithmic bytes): same ceiling as G1, the L2 reduction rate on a 7000-cell tally",
         "g4": _STREAMED % "720 000 cells: kappa_factor, tallies, xT_ech = 17 MB, L2-resident",
         "g5": _STREAMED % "997 016 cells: seeds, neighbour lists (15.5 per cell), kappa_factor, tallies = 110 MB"}
_ZERO = "tallies are re-zeroed (memset) every step; "
L2_POLICY = {"g1": _ZERO + "working set is L2-resident by design (0.5 MB tables)", "g2": _ZERO + "working set is L2-resident by design (0.5 MB tables)",
             "g3": _ZERO + "per-cell tables (300 MB) are larger than L2", "g4": _ZERO + "per-cell data 17 MB, L2-resident; every packet takes its own path",
             "g5": _ZERO + "mesh and per-cell data (110 MB) are of the order of L2; every packet takes its own path"}


_REAL_STDOUT = None


def emit(line):
    """the one JSON line, on the process's real stdout (see main)"""
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.buffer.write(data); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def make_problem(n2_total, walker_factory=None, workload="g1"):
    from mcfost_b200 import synthetic as S
    if workload == "g2":
        return S.pascucci_like(tau_V=100.0, n_photons_eq_th=n2_total)
    if workload == "g3":
        return S.ref41_multi_like(n_photons_eq_th=n2_total, n_rad=100, nz=70, n_rad_in=20, tau_mid=1.0e5)
    if workload == "g4":
        return S.ref41_3d_like(n_photons_eq_th=n2_total, tau_mid=1.0e3, n_rad=100, nz=50, n_az=72)
    if workload == "g5":
        return S.voronoi_sph_disk(n_points=1000000, n_photons_eq_th=n2_total, tau_mid=1.0e3,
                                  cache=os.path.join(ROOT, "data_cache", "g5_1000000.npz"))
    P = S.ref41_like(n_photons_eq_th=n2_total, dark_zone=False)      # L_packet_th = L_tot / (128 * n2_total)
    if walker_factory is not None:
        P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, walker_factory(P))
        S.repartition_energie(P)
    return P


def set_budget(P, n2_total):
    from mcfost_b200 import synthetic as S
    P.n_photons_eq_th = n2_total
    S.repartition_energie(P)                   # same model, L_packet_th for this packet count


def host_threads(O):
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # turn the OpenMP baseline into a single-thread run, so the affinity mask decides, not the environment
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    return max(O.lib.oracle_max_threads(), avail)


def cpu_run(O, n2, nthr, mrw):
    """oracle-OpenMP (reference restatement), timing flavour; returns (packets, seconds)."""
    t0 = time.perf_counter()
    t = O.run(n_threads=nthr, n_photons2=n2, lMRW=mrw, **FLAGS)
    return float(t.stats[0]), time.perf_counter() - t0


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding
    try:
        binding.build(fast_native=True)      # -march=native on the box that runs it
    except Exception:
        binding.build()
    from oracle.binding import Oracle
    n2 = args.cpu_n2
    P = make_problem(n2, lambda P: Oracle(P).dark_zone_walker(), args.workload)
    FLAGS.update(WORKLOADS[args.workload]["flags"])
    O = Oracle(P, fast=True)
    nthr = args.cpu_threads or host_threads(O)
    cpu_run(O, max(1, n2 // 50), nthr, args.mrw)           # warm-up (thread pool, page faults)
    times, packets = [], 0.0
    for i in range(args.warmup + args.steps):
        pk, dt = cpu_run(O, n2, nthr, args.mrw)
        if i >= args.warmup:
            times.append(dt); packets += pk
    total = sum(times)
    val = packets / total
    # the budgets of the GPU arm's sweep, same packet counts (the largest is bounded by --cpu-sweep-max)
    sweep = []
    for s2 in (SWEEP_N2 if args.workload == "g1" else ()):
        if 128 * s2 > args.cpu_sweep_max:
            sweep.append({"packets": 128 * s2, "value": None, "note": "skipped: above --cpu-sweep-max (host time)"})
            continue
        set_budget(P, s2); O.set_emission(P)
        pk, dt = cpu_run(O, s2, nthr, args.mrw)
        sweep.append({"packets": int(pk), "value": pk / dt, "ms": 1e3 * dt})
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["name"], "packets_per_step": int(128 * n2), "lMRW": int(args.mrw)},
            "sweep": sweep,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthr, "kind": "port",
                             "sample": f"oracle-OpenMP (reference restatement, the Fortran cannot be built here), {128 * n2} packets per step, schedule(dynamic,1) over 128 chunks"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from mcfost_b200 import api, synthetic as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the photon loop has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    W = WORKLOADS[args.workload]
    FLAGS.update(W["flags"])
    P = make_problem(args.n2 * world, workload=args.workload)
    loop = api.PhotonLoop(P, device=local, rank=rank, n_ranks=world)
    if args.workload == "g1":
        # dark zone by the library (mcfost_b200_define_dark_zone, optical_depth.f90:1425-1651), which also installs it
        P.l_dark_zone = loop.define_dark_zone(P.lambda_seuil, 1500.0, P.r_grid, P.z_grid, [(1, P.n_rad)])["l_dark_zone"]
        S.repartition_energie(P)
        loop.upload_dark_zone(P.l_dark_zone)
        loop.upload_emission(P)
    dev = torch.device("cuda", local)
    stream = torch.cuda.ExternalStream(loop.stream(), device=dev)
    flags = dict(FLAGS, lMRW=int(args.mrw))
    view = [None]

    def reduce_tallies():
        if world > 1:
            if view[0] is None:
                v64, _ = loop.tally_buffers()
                view[0] = torch.as_tensor(v64, device=dev)
            with torch.cuda.stream(stream):
                dist.all_reduce(view[0], op=dist.ReduceOp.SUM)     # one NCCL all-reduce per call

    def step(n2, call_index):
        """one blocking call: launch, (all-reduce,) wait"""
        loop.launch(1, 1, n2, 1.0e30, 1, call_index=call_index, reset_tallies=1, **flags)
        reduce_tallies()
        loop.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        loop.sync()

    def timed(n2, k, first_index):
        """k blocking calls back to back: device time (CUDA events on the library's stream), max over ranks"""
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for i in range(k):
            step(n2, first_index + i)
        ev1.record(stream)
        barrier()
        tm = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm[0])

    n2 = args.n2 * world                      # weak scaling: 128/world chunks x (n2*world) packets per rank
    for i in range(args.warmup):
        step(n2, i)
    sampler = ClockSampler(local)
    sampler.start()
    dev_ms = timed(n2, args.steps, args.warmup)
    sampler.stop_flag = True
    d_last = loop.debug_counters()
    launches_per_step = d_last["launches"]      # counted by the library: packet-per-lane kernel + counter hand-over + packet-per-warp kernel (small budgets: the latter alone)
    t_last = loop.download(want_xI=False)
    stats = t_last.stats.copy()                # whole-job counts after the all-reduce
    packets_per_step = 128 * args.n2 * world
    value = packets_per_step * args.steps / (dev_ms * 1e-3)
    last_ms = dev_ms / args.steps

    # ---- e2e: the same blocking calls through the C ABI with HOST buffers inside the timed region.  Every step uploads
    # its emission tables from pinned host memory (repartition_energie output changes every temperature iteration:
    # mcfost_b200_upload_emission), runs (mcfost_b200_launch + all-reduce + mcfost_b200_sync) and brings its tallies back
    # to host arrays (mcfost_b200_download).  Nothing is created on the device, nothing overlaps.
    pinned_keep = []

    def pin(a, dtype):
        a = np.asfortranarray(np.asarray(a, dtype=dtype))
        try:
            t_ = torch.from_numpy(np.ascontiguousarray(a.T)).pin_memory()      # a.T of an F-ordered array is C-contiguous
        except Exception:
            return a
        pinned_keep.append(t_)
        return t_.numpy().T                                                     # F-ordered view of the pinned block
    for nm in ("spectre_emission_cumul", "frac_E_stars", "frac_E_disk", "prob_E_cell"):
        setattr(P, nm, pin(getattr(P, nm), np.float64))
    P.CDF_E_star = pin(P.CDF_E_star, np.float32)
    tally_names = ("xKJ_abs", "xT_ech", "n_phot_envoyes", "sed", "sed_q", "sed_u", "sed_v", "n_phot_sed",
                   "sed_star", "sed_star_scat", "sed_disk", "sed_disk_scat", "stats")

    def e2e_call(n2_, call_index):
        loop.upload_emission(P)
        h2d_ = sum(a.nbytes for a in loop._e.keep.values())
        step(n2_, call_index)
        t = loop.download(want_xI=False)
        d2h_ = sum(getattr(t, nm).nbytes for nm in tally_names)
        return h2d_, d2h_

    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for i in range(args.steps):
        h2d, d2h = e2e_call(n2, 1000 + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = packets_per_step * args.steps / float(te[0])

    # ---- strong scaling (N > 1): the fixed 1.28e8-packet job over the N ranks (128/N chunks x n2 packets per rank)
    strong = None
    if world > 1:
        set_budget(P, args.n2); loop.upload_emission(P)
        step(args.n2, 3000)
        s_ms = timed(args.n2, args.strong_steps, 3001)
        strong = {"scaling": "strong", "packets_per_step": int(128 * args.n2), "steps": args.strong_steps,
                  "ms_per_step": s_ms / args.strong_steps, "value": 128 * args.n2 * args.strong_steps / (s_ms * 1e-3), "unit": UNIT}
        set_budget(P, args.n2 * world)

    # ---- budget sweep (N = 1): one blocking e2e call per budget, host buffers, best of 2 after one warm-up call
    sweep = None
    if world == 1 and not args.no_sweep and args.workload == "g1":
        sweep = []
        for s2 in SWEEP_N2:
            set_budget(P, s2)
            e2e_call(s2, 4000)
            best, best_dev = 1e30, 0.0
            for rep in range(2):
                t1 = time.perf_counter()
                e2e_call(s2, 4001 + rep)
                dt = time.perf_counter() - t1
                if dt < best:
                    best, best_dev = dt, loop.last_kernel_ms()
            sweep.append({"packets": 128 * s2, "value": 128 * s2 / best, "ms": 1e3 * best, "device_ms": best_dev})
        set_budget(P, args.n2); loop.upload_emission(P)

    if rank == 0:
        peak, peak_src = peaks()
        nb = stats[1] * W["b_step"] + stats[3] * B_SCA + stats[4] * B_ABS + stats[0] * b_packet(P.n_cells)
        nb_per_gpu = nb / world
        hbm_achieved = nb_per_gpu / (last_ms * 1e-3) / 1e9
        atomics_per_s = stats[1] / world / (last_ms * 1e-3)            # one fp64 reduction per crossed cell
        l2 = _json("r02_l2_atomic.json")
        traffic = (_json("r02_traffic.json") or {}).get(args.workload)
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import binding
            try:
                binding.build(fast_native=True)
            except Exception:
                binding.build()
            from oracle.binding import Oracle
            import copy
            Pc = copy.copy(P)
            set_budget(Pc, args.cpu_n2)
            O = Oracle(Pc, fast=True)
            nthr = host_threads(O)
            cpu_run(O, max(1, args.cpu_n2 // 50), nthr, args.mrw)
            pk, dt = cpu_run(O, args.cpu_n2, nthr, args.mrw)
            cpu = {"value": pk / dt, "unit": UNIT, "cores": nthr, "kind": "port",
                   "sample": f"oracle-OpenMP (reference restatement), same model, {int(pk)} packets in {dt:.2f} s wall on {nthr} threads"}
        if l2 and "g1_hits" in l2 and args.workload in ("g1", "g2", "g3"):
            l2_peak = float(l2["g1_hits"]["red_f64_per_s"])
            roof = {"bound": "l2_atomic", "achieved": atomics_per_s * 8e-9, "peak": l2_peak * 8e-9, "unit": "GB/s", "frac": atomics_per_s / l2_peak,
                    "peak_source": "measured: red.global.add.f64 into 7000 L2-resident doubles with G1's per-cell crossing distribution (tools/l2_atomic_peak.cu, profiles/r02_l2_atomic.json); payload bytes of the reductions"}
        else:
            roof = {"bound": "hbm", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak, "peak_source": peak_src}
        roof.update({"traffic": (traffic or {}).get("dram_bytes_per_launch"),
                     "algorithmic_bytes_per_launch": nb_per_gpu, "hbm_algorithmic_GBps": hbm_achieved, "hbm_frac": hbm_achieved / peak,
                     "kernel": KERNELS[args.workload],
                     "kernel_ms": last_ms,
                     "phases_ms_last_call": {"packet_per_lane_dry": d_last["steady_ms"], "packet_per_lane_end": d_last["main_end_ms"],
                                             "stragglers_start": d_last["straggler_start_ms"], "stragglers_end": d_last["straggler_end_ms"], "parked": d_last["parked"]},
                     "note": NOTES[args.workload],
                     "fp64_reductions_per_s": atomics_per_s, "steps_per_s": stats[1] / world / (last_ms * 1e-3), "interactions_per_s": stats[2] / world / (last_ms * 1e-3),
                     "steps_per_packet": stats[1] / stats[0], "interactions_per_packet": stats[2] / stats[0], "mrw_steps_per_packet": stats[9] / stats[0]})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": W["name"], "packets_per_step": int(packets_per_step), "lMRW": int(args.mrw),
                           "parallelism": f"packets x{world} (replicated grid, 1 all-reduce/step)",
                           "calls": "one blocking call per step (launch + all-reduce + sync), nothing overlapped",
                           "l2_policy": L2_POLICY[args.workload]},
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "host_buffers": "emission tables in pinned host memory (%d pinned arrays), tallies into host numpy arrays" % len(pinned_keep)},
                "gpu_launches": int(args.steps * (launches_per_step + 1)),   # + fill_int_kernel (xT_ech reset)
                "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu}
        if sweep is not None:
            line["sweep"] = sweep
        if strong is not None:
            line["strong"] = strong
        emit(line)
    loop.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n2", type=int, default=0, help="packets per chunk per GPU (128 chunks); default: 1000000 for g1 (1.28e8 packets per step per GPU), "
                    "10000 for g2 (the Pascucci_3.0.para budget), 100000 for g3 / g4, 20000 for g5")
    ap.add_argument("--cpu-n2", type=int, default=8000, help="packets per chunk of the bounded CPU sample")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--cpu-sweep-max", type=float, default=1.28e7, help="largest sweep budget the reference arm runs (1.28e8 takes minutes of host time)")
    ap.add_argument("--mrw", type=int, default=0, help="1: modified random walk on (both arms); the reference's own behaviour is off")
    ap.add_argument("--strong-steps", type=int, default=3)
    ap.add_argument("--workload", default="g1", choices=sorted(WORKLOADS), help="g1 = the headline (ref4.1-like); g2 / g3 / g4 / g5: the other BASELINE configs")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    args = ap.parse_args()
    # the contract is ONE JSON line on stdout: libraries that write there themselves (NCCL prints its version line to
    # stdout when the first communicator is created) are sent to stderr; the JSON line goes to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.n2 <= 0:
        args.n2 = WORKLOADS[args.workload]["n2"]
    args.cpu_n2 = min(args.cpu_n2, args.n2)
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
