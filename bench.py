#!/usr/bin/env python
"""Benchmark of the photon-packet loop (BASELINE.json metric: photon packets/sec on the
ref4.1 disk at 1/2/4/8 B200 vs host OpenMP).

    python bench.py --gpus N --steps K --warmup W            # CUDA arm
    python bench.py --impl reference --gpus N --steps K ...   # host-core OpenMP arm

One "step" = one thermal-mode pass of mc_photon_loop (dust_transfer.f90:439) over one batch of
packets on the synthetic ref4.1-like model G1: 128 chunks x --n2 packets per GPU
(weak scaling: chunks are dealt round-robin to ranks, n2 is scaled by the number of ranks).
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
# SURVEY 8d algorithmic bytes: per cell-crossing step (2D cyl), per scattering, per absorption, per packet
B_STEP, B_SCA, B_ABS = 112.0, 72.0, 112.0


def b_packet(n_cells):
    return np.ceil(np.log2(n_cells)) * 8.0 + 96.0


def measured_traffic(packets_per_launch):
    """DRAM bytes per launch of the photon-loop kernel from the committed ncu launch list of this same
    command (profiles/r01_traffic.json); None when it was taken at another packet budget."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        if int(t["packets_per_launch"]) == int(packets_per_launch):
            return float(t["dram_bytes_per_launch_mean"])
    except Exception:
        pass
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


def make_problem(args, walker_factory=None, world=1):
    from mcfost_b200 import synthetic as S
    P = S.ref41_like(n_photons_eq_th=args.n2 * world, dark_zone=False)      # L_packet_th = L_tot / (128 * n2 * world)
    if walker_factory is not None:
        P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, walker_factory(P))
        S.repartition_energie(P)
    return P


def cpu_run(P, n2, threads=0, rank=0, n_ranks=1):
    """oracle-OpenMP (reference restatement), timing flavour; returns (packets, seconds, stats, threads)."""
    from oracle import binding
    from oracle.binding import Oracle
    O = Oracle(P, fast=True)
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # turn the OpenMP baseline into a single-thread run, so the affinity mask decides, not the environment
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    nthr = threads or max(O.lib.oracle_max_threads(), avail)
    O.run(n_threads=nthr, n_photons2=max(1, n2 // 50), **FLAGS)           # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    t = O.run(n_threads=nthr, n_photons2=n2, **FLAGS)
    dt = time.perf_counter() - t0
    return float(t.stats[0]), dt, t.stats.copy(), nthr


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
    args.n2 = args.cpu_n2                      # the sample's own packet count sets L_packet_th
    P = make_problem(args, lambda P: Oracle(P).dark_zone_walker())
    n2 = args.cpu_n2
    times, packets = [], 0.0
    for i in range(args.warmup + args.steps):
        pk, dt, st, nthr = cpu_run(P, n2)
        if i >= args.warmup:
            times.append(dt); packets += pk
    total = sum(times)
    val = packets / total
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "G1 ref4.1-like: cylindrical 100x70x1, 50 lambda, n_T=100, thermal step, tau_mid(0.81um)=1e5, dark zone tau>1500",
                       "packets_per_step": int(128 * n2)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthr, "kind": "port",
                             "sample": f"oracle-OpenMP (reference restatement, the Fortran cannot be built here), {128 * n2} packets per step, schedule(dynamic,1) over 128 chunks"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


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

    P = make_problem(args, world=world)
    loop = api.PhotonLoop(P, device=local, rank=rank, n_ranks=world)
    # dark zone via the library's own deterministic ray-walk kernel (define_dark_zone step 4)
    P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, loop.dark_zone_walker())
    S.repartition_energie(P)
    loop.upload_dark_zone(P.l_dark_zone)
    loop.upload_emission(P)
    # Two handles (own stream, own tallies, own constant bank) are used alternately so that the
    # drain-out of step i (a few very long-lived packets, most SMs already free) overlaps the start of
    # step i+1.  Every step is a complete, separately reduced mc_photon_loop call.
    loops = [loop] + [api.PhotonLoop(P, device=local, rank=rank, n_ranks=world) for _ in range(max(1, args.pipeline) - 1)]
    if len(loops) > 1 and args.overlap_sms > 0:
        for l in loops:
            l.set_overlap(args.overlap_sms, max(1, args.overlap_sms // (len(loops) - 1)))   # main launches leave these SMs to the straggler launches of the other handles
    dev = torch.device("cuda", local)
    n2 = args.n2 * world                      # weak scaling: 128/world chunks x (n2*world) packets per rank
    streams = [torch.cuda.ExternalStream(l.stream(), device=dev) for l in loops]
    views = [None] * len(loops)

    def step(i, call_index):
        k = i % len(loops)
        loops[k].launch(1, 1, n2, 1.0e30, 1, call_index=call_index, reset_tallies=1, **FLAGS)
        if world > 1:
            if views[k] is None:
                v64, _ = loops[k].tally_buffers()
                views[k] = torch.as_tensor(v64, device=dev)
            with torch.cuda.stream(streams[k]):
                dist.all_reduce(views[k], op=dist.ReduceOp.SUM)     # one NCCL all-reduce per call

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for l in loops:
            l.sync()

    for i in range(args.warmup):
        step(i, i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in loops]
    ev0.record(streams[0])
    for s_ in streams[1:]:
        s_.wait_event(ev0)
    for i in range(args.steps):
        step(i, args.warmup + i)
    for e_, s_ in zip(ends, streams):
        e_.record(s_)
    barrier()
    sampler.stop_flag = True
    dev_ms = max(ev0.elapsed_time(e_) for e_ in ends)
    # stats of the last step (whole-job counts after the all-reduce), for the roofline
    t_last = loops[(args.steps - 1) % len(loops)].download(want_xI=False)
    stats = t_last.stats.copy()
    tm = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dev_ms = float(tm[0])
    last_ms = dev_ms / args.steps             # average device time per launch over the timed region
    packets_per_step = 128 * args.n2 * world          # whole job
    value = packets_per_step * args.steps / (dev_ms * 1e-3)

    # ---- e2e: the same calls through the C ABI with HOST buffers inside the timed region.  Every step uploads its
    # emission tables from host memory (repartition_energie output changes every temperature iteration / wavelength:
    # mcfost_b200_upload_emission), launches (mcfost_b200_launch), and its tallies are brought back to host arrays
    # (mcfost_b200_sync + mcfost_b200_download).  The handles are used in turn exactly as in the device-timed loop, so
    # the download of step i overlaps the run of step i+1; nothing is created on the device.
    e2e_steps = args.steps
    # the host copies of the emission tables live in PINNED memory for the e2e loop (the ctypes layer passes the
    # arrays through untouched when dtype and Fortran layout already match)
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
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    pending = [None] * len(loops)
    for i in range(e2e_steps + len(loops)):
        k = i % len(loops)
        if pending[k] is not None:
            loops[k].sync()
            t = loops[k].download(pending[k], want_xI=False)
            d2h = sum(getattr(t, nm).nbytes for nm in ("xKJ_abs", "xT_ech", "n_phot_envoyes", "sed", "sed_q", "sed_u", "sed_v", "n_phot_sed",
                                                       "sed_star", "sed_star_scat", "sed_disk", "sed_disk_scat", "stats"))
            pending[k] = None
        if i < e2e_steps:
            loops[k].upload_emission(P)
            h2d = sum(a.nbytes for a in loops[k]._e.keep.values())
            step(i, 1000 + i)
            pending[k] = loops[k]._last_run
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = packets_per_step * e2e_steps / float(te[0])

    if rank == 0:
        peak, peak_src = peaks()
        nb = stats[1] * B_STEP + stats[3] * B_SCA + stats[4] * B_ABS + stats[0] * b_packet(P.n_cells)
        nb_per_gpu = nb / world
        achieved = nb_per_gpu / (last_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import binding
            try:
                binding.build(fast_native=True)
            except Exception:
                binding.build()
            import copy
            Pc = copy.copy(P)
            Pc.n_photons_eq_th = args.cpu_n2
            S.repartition_energie(Pc)                 # same model, L_packet_th for the sample's packet count
            pk, dt, st, nthr = cpu_run(Pc, args.cpu_n2)
            cpu = {"value": pk / dt, "unit": UNIT, "cores": nthr, "kind": "port",
                   "sample": f"oracle-OpenMP (reference restatement), same model, {int(pk)} packets in {dt:.2f} s wall on {nthr} threads"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "G1 ref4.1-like: cylindrical 100x70x1, 50 lambda, n_T=100, thermal step, tau_mid(0.81um)=1e5, dark zone tau>1500",
                           "packets_per_step": int(packets_per_step), "parallelism": f"packets x{world} (replicated grid, 1 all-reduce/step)", "pipeline": f"{len(loops)} handles alternate so that the drain-out of one step overlaps the next" + (f"; {args.overlap_sms} SMs reserved for the straggler launches" if len(loops) > 1 and args.overlap_sms > 0 else ""),
                           "l2_policy": "tallies are re-zeroed (memset) every step; working set is L2-resident by design (0.5 MB tables)"},
                "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "host_buffers": "emission tables in pinned host memory (%d pinned arrays), tallies into host numpy arrays" % len(pinned_keep)},
                "gpu_launches": int(args.steps * (3 if (len(loops) > 1 and args.overlap_sms > 0) else 2)),   # per step: mc_photon_loop_kernel (+ its straggler launch) + fill_int_kernel (xT_ech reset)
                "clocks": sampler.summary(),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(packets_per_step / world),
                             "algorithmic_bytes_per_launch": nb_per_gpu,
                             "peak_source": peak_src, "kernel": "mc_photon_loop_kernel<GeomCyl<0,1>,1,BANK,0> (main + straggler launch)", "kernel_ms": last_ms,
                             "launch_ms_last_call": loops[(args.steps - 1) % len(loops)].last_kernel_ms(),
                             "note": "algorithmic bytes (SURVEY 8d) of one call / device time per call over the timed region (calls on the handles overlap, so one call's own launches last longer: launch_ms_last_call); tables are L2-resident so the path is latency/issue-bound, not HBM-bound",
                             "steps_per_s": stats[1] / world / (last_ms * 1e-3), "interactions_per_s": stats[2] / world / (last_ms * 1e-3),
                             "steps_per_packet": stats[1] / stats[0], "interactions_per_packet": stats[2] / stats[0]},
                "cpu_baseline": cpu}
        print(json.dumps(line))
    for l in loops:
        l.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n2", type=int, default=1000000, help="packets per chunk per GPU (128 chunks): 1.28e8 packets per step per GPU")
    ap.add_argument("--cpu-n2", type=int, default=8000, help="packets per chunk of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--pipeline", type=int, default=2, help="handles used alternately (1 = strictly serial steps)")
    ap.add_argument("--overlap-sms", type=int, default=16, help="SMs reserved for straggler launches when pipelining (0 = off)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
