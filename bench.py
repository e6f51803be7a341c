#!/usr/bin/env python3
"""bench.py -- particle-steps/s of the per-timestep PIC loop (deposit + solve + push) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload gauss_fp|ngp|2d3v] [--impl reference]

Headline workload (BASELINE.json configs[3]): scaled synthetic Gaussian fixed-point two-stream 1D1V,
N=4096 cells, 2^28 particles per GPU (weak scaling), src/GaussianFixedPointQuiet.jl loop body with the
+-6 stencil and l=1e-8, seeded-uniform two-stream start (the GaussianFixedPoint.jl x=rand(P) pattern).
A "step" is one PIC time step = S fixed-point sweeps (S reported).  One JSON line on stdout.

Timing: W untimed warm-up steps, then exactly K steps bracketed by barrier + synchronize; CUDA events
on the library's own stream; max over ranks.  Inputs (8 GiB of particle state per GPU) exceed L2.
`value` has the particle state resident in HBM; `e2e` pushes the whole particle state host->device and
device->host through the C ABI every step (picgolf_step_streamed, pinned host buffers, copies overlapped
with the steps of the neighbouring calls); the serial set/step/get sequence is reported next to it.
"""
import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec (deposit+solve+push)"
UNIT = "particle-steps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled from before the warm-up to the end of the timed region (NVML from a thread, as
    fast as NVML answers; nvidia-smi -lms 100 as the fallback).  Every sample carries a time stamp, so the report gives the
    samples that fell INSIDE the timed region and, because the driver's 20-step run lasts only ~0.1 s, also everything
    sampled under load since the warm-up began."""
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu):
        self.gpu, self.rows, self.mx, self.stop_flag, self.t, self.proc, self.how = gpu, [], None, False, None, None, None

    def _nvml_loop(self, h, nv):
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((time.perf_counter(), sm, frozenset(n for n, b in self.BITS.items() if r & b)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    pass
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.how = "nvml thread"
            self.t = threading.Thread(target=self._nvml_loop, args=(h, nv), daemon=True)
            self.t.start()
            return
        except Exception:
            self.how = "nvidia-smi -lms 100"
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                for l in self.proc.stdout:
                    f = [x.strip() for x in l.split(",")]
                    try:
                        self.mx = float(f[1])
                        self.rows.append((time.perf_counter(), float(f[0]), frozenset(
                            n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]) if v.lower().startswith("active"))))
                    except (ValueError, IndexError):
                        continue
            self.t = threading.Thread(target=pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
        if self.t:
            self.t.join(timeout=2)
        rows = list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0, "how": self.how}
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        use = inside if len(inside) >= 3 else rows
        reasons = sorted(set().union(*[r[2] for r in use]))
        return {"sm_mhz": float(np.median([r[1] for r in use])), "sm_max_mhz": self.mx, "reasons": reasons, "samples": len(use),
                "samples_inside_timed_region": len(inside), "samples_since_warmup_began": len(rows),
                "window": "timed region" if use is inside else "warm-up + timed region (under load throughout; the timed region alone held fewer than 3 samples)",
                "how": self.how}


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def make_sim(pg, workload, per_gpu, rank, world, device, T, deposit_mode=0):
    P = per_gpu * world
    if workload == "gauss_fp":
        sim = pg.gaussian_fixed_point(N=4096, P=P, T=T, W=400.0, l=1e-8, half_width=6, rank=rank, nranks=world, device=device,
                                      deposit_mode=deposit_mode)
        name = f"config4 scaled Gaussian fixed-point two-stream 1D1V N=4096 P={P} (2^{int(math.log2(per_gpu))}/GPU) +-6 l=1e-8 seeded-uniform start"
        bytes_per_unit = None  # 32*(S+1), needs the measured sweep count
    elif workload == "ngp":
        sim = pg.ngp_fourier(N=4096, P=P, NT=T, W=256.0, rank=rank, nranks=world, device=device)
        name = f"config1-scaled NGP leapfrog 1D1V N=4096 P={P} seeded-uniform two-stream"
        bytes_per_unit = 32.0
    elif workload == "2d3v":
        sim = pg.electrostatic_2d3v(NX=256, NY=256, P=P, T=T, NS=1, rank=rank, nranks=world, device=device)
        name = f"config5 Electrostatic2D3V CIC+Boris 256x256 P={P} Maxwellian start"
        bytes_per_unit = 80.0
    else:
        raise SystemExit(f"unknown workload {workload}")
    return sim, name, bytes_per_unit


def init_sim(sim, workload, seed=1234, vth=None):
    sim.init_synthetic(seed=seed, vth=getattr(sim, "vth", 0.0) if vth is None else vth)


def traffic_of(key):
    """Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and FP64 instructions per particle of the dominant
    kernel, from the committed `ncu --set full` capture of this kernel at 2^28 particles (ncu cannot run inside a timed bench;
    the file names its capture)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))[key]
            t = dict(t)
            t["source"] = "profiles/" + name
            return t
        except Exception:
            continue
    return None


def bind_to_gpu_numa(local):
    """Pin this rank to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE any pinned host buffer
    is allocated, so that the e2e arm's staging memory is first-touched on the GPU's own NUMA node and the eight ranks of a box
    do not all stream through one socket.  Returns a short description for the JSON line; harmless where it cannot apply."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = open(dev + "/numa_node").read().strip()
        cpus = set()
        for part in open(dev + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"numa_node": node, "bound": False, "why": "no local CPU in this process's affinity mask"}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "bound": True, "cpus": len(cpus)}
    except Exception as e:  # no NVML / no sysfs entry: leave the affinity alone
        return {"bound": False, "why": type(e).__name__}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import particleincellcodegolf.jl_b200 as pg
    from particleincellcodegolf.jl_b200 import distributed as pgd

    rank, world, local = pgd.env_rank()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    # several ranks on one box: each next to its own GPU (a single rank keeps all host cores for the CPU baseline's threads)
    numa = bind_to_gpu_numa(local) if world > 1 else {"bound": False, "why": "single rank"}
    torch.cuda.set_device(local)
    sampler = ClockSampler(local)
    sampler.start()  # before the warm-up: the timed region of the driver's run is only ~0.1 s long
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, Wm = args.steps, max(args.warmup, 3)
    per_gpu = 1 << args.log2_particles_per_gpu
    sim, name, bpu = make_sim(pg, args.workload, per_gpu, rank, world, local, T=K + Wm + 64,
                              deposit_mode={"auto": 0, "atomic": 1, "sorted": 2, "poly": 3}[args.deposit_mode])
    if world > 1:
        pgd.connect(sim)
    sim_peer = world > 1 and sim.peer_status[0]
    init_sim(sim, args.workload)
    stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local))

    def barrier():
        sim.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident timing -------------------------------------------------------------
    sampler.rows.clear()  # samples from here on: warm-up and timed region, both under load
    sim.step(Wm)
    barrier()
    sorts0 = sim.sort_stats()[0]
    l0 = sim.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record(stream)
    sim.step(K)  # one C-ABI call enqueues all K steps (no host sync inside)
    e1.record(stream)
    barrier()
    tw1 = time.perf_counter()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(tw0, tw1)
    launches = sim.launches - l0
    sorts_timed = sim.sort_stats()[0] - sorts0
    fused_main = sim.fused_sorts if args.workload == "gauss_fp" else 0
    P = sim.cfg.P
    value = P * K / (ms * 1e-3)
    D, sw = sim.diagnostics()
    sweeps = sw[Wm:Wm + K].astype(float) if args.workload == "gauss_fp" else np.ones(K)
    mean_sweeps = float(sweeps.mean())

    # ---- roofline of the dominant kernel (the particle pass), timed live with CUDA events ------
    # (stage timers on: the fixed launch schedule, whose events bracket every kernel -- the kernels are those of the timed region)
    sim.stage_timing(True)
    sim.stage_times(reset=True)
    poly = args.workload == "gauss_fp" and sim.deposit_path == pg.DEPOSIT_POLY
    nroof = 40 if poly else 12  # long enough to contain a re-sort
    s0 = sim.sort_stats()[0]
    sim.step(nroof)
    st = sim.stage_times(reset=True)
    sim.stage_timing(False)
    _, sw2 = sim.diagnostics()
    if args.workload == "gauss_fp":
        passes = float(sw2[Wm + K:Wm + K + nroof].sum())  # S solves -> S particle passes per step (the k=0 pass is fused into the previous step's final pass)
        alg_bytes_launch = 32.0 * per_gpu
        kernel = ("fp_pass_poly<FIRST> (sub-cell gather polynomial + implicit-midpoint update + register moment deposit)" if poly else
                  "fp_pass_sorted<FIRST,NP=2> (gather + implicit-midpoint update + deposit; 3 variants: first/middle/final pass)")
        tkey = "gauss_fp_poly" if poly else "gauss_fp"
    elif args.workload == "ngp":
        passes = float(nroof)  # the warm-up call left the first step's charge deposited: K steps are K passes
        alg_bytes_launch = 32.0 * per_gpu
        kernel = "lf_pass_ngp_tma (TMA-staged tiles: drift + kick + drift + NGP deposit)"
        tkey = "ngp"
    else:
        passes = float(nroof)
        alg_bytes_launch = 80.0 * per_gpu
        kv = os.environ.get("PICGOLF_2D_KERNEL", "stream")
        kernel = "particles_2d3v_" + ("stream<4,8,512>" if kv == "stream" else kv) + " (gather + boris + move + CIC deposit)"
        tkey = "2d3v"
    launch_ms = st["particles"] / passes
    peak, peak_src = peaks()
    achieved = alg_bytes_launch / (launch_ms * 1e-3) / 1e9
    tj = traffic_of(tkey) if per_gpu == 1 << 28 else None  # the capture is of a 2^28-particle launch
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": tj["traffic_bytes"] if tj else None,
                "traffic_source": (tj["source"] + ": " + tj.get("capture", "")) if tj else None,
                "launch_ms": launch_ms, "algorithmic_bytes_per_launch": alg_bytes_launch,
                "stage_ms_per_step": {k: v / nroof for k, v in st.items()},
                "resorts_in_stage_window": sim.sort_stats()[0] - s0, "stage_window_steps": nroof}

    # ---- FP64 pipe view of the same kernel (the second bound of the erf-shape passes) ----------
    fp64 = None
    if args.workload == "gauss_fp":
        peak_tf = pg.fp64_peak_tflops()
        per_pass = (tj or {}).get("fp64_inst_per_particle_pass", 47 if poly else 255)  # ncu: smsp__inst_executed_pipe_fp64.sum * 32 / particles
        fp64_inst = float(sum(per_pass * int(s_) for s_ in sw2[Wm + K:Wm + K + nroof])) * per_gpu
        ach = fp64_inst * 2.0 / (st["particles"] * 1e-3) / 1e12  # counted as 2 flops per lane-instruction (FMA-equivalent)
        fp64 = {"measured_peak_tflops": peak_tf, "achieved_tflops_fma_equiv": ach, "frac": ach / peak_tf,
                "fp64_instructions_per_particle_pass": per_pass,
                "note": "FP64 lane-instructions of the pass kernels (ncu count, profiles/) x2 / kernel time measured here"}

    # ---- end to end through the C ABI with host buffers ------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, pg, sim, torch, barrier, max_over_ranks, stream, P, world)
        e2e["host_affinity"] = numa

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) -------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args.workload, args.cpu_log2_particles, args.cpu_steps)

    # ---- the same workload in the warm (saturated two-stream) regime ------------------------------------------
    warm = None
    if args.warm and args.workload == "gauss_fp":
        sim.close()
        sim, _, _ = make_sim(pg, "gauss_fp", per_gpu, rank, world, local, T=128, deposit_mode=0)
        if world > 1:
            pgd.connect(sim)
        stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local))
        warm = {}
        Ww = 40
        for vth in (0.05, 0.3, 1.0):
            init_sim(sim, "gauss_fp", seed=99, vth=vth)
            sim.step(Ww)  # long enough for the adaptive re-sort interval to settle (the flush probe lags 4 steps)
            barrier()
            so, lo_, fo = sim.sort_stats()[0], sim.launches, sim.fused_sorts
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            Kw = 24
            g0.record(stream)
            sim.step(Kw)
            g1.record(stream)
            barrier()
            msw = max_over_ranks(g0.elapsed_time(g1))
            sww = sim.diagnostics()[1][Ww:Ww + Kw].astype(float)
            warm[f"vth={vth}"] = {"value": P * Kw / (msw * 1e-3), "unit": UNIT, "ms_per_step": msw / Kw, "steps": Kw, "warmup": Ww, "mean_sweeps_per_step": float(sww.mean()),
                                  "resorts": sim.sort_stats()[0] - so, "resorts_fused_into_the_passes": sim.fused_sorts - fo, "hbm_roofline_frac_step": (P / world) * Kw * 32.0 * float(sww.mean()) / (msw * 1e-3) / (peak * 1e9),
                                  "start": f"seeded-uniform x, v = +-1 + {vth}*N(0,1): beams as warm as after saturation (vortices), the bins of the sorted order shear apart within a few steps"}
        warm["note"] = ("the headline is measured in the cold-beam phase (bins drift rigidly, one re-sort per 16-64 steps); with warm beams the flush "
                        "probe shortens the interval, down to a re-sort fused into the passes of every step (counted here)")

    # ---- the other two scoped workloads, briefly (device-resident, same timing rules) ----------------------
    others = None
    if args.others and args.workload == "gauss_fp":
        sim.close()
        others = {}
        for wl in ("ngp", "2d3v"):
            s2, name2, bpu2 = make_sim(pg, wl, per_gpu, rank, world, local, T=64)
            if world > 1:
                pgd.connect(s2)
            init_sim(s2, wl)
            st2 = torch.cuda.ExternalStream(s2.stream, device=torch.device("cuda", local))
            s2.step(4)
            s2.synchronize(); torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            Ko = 16
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(st2)
            s2.step(Ko)
            g1.record(st2)
            s2.synchronize(); torch.cuda.synchronize()
            mso = max_over_ranks(g0.elapsed_time(g1))
            vo = s2.cfg.P * Ko / (mso * 1e-3)
            others[wl] = {"workload": name2, "value": vo, "unit": UNIT, "steps": Ko, "ms_per_step": mso / Ko,
                          "algorithmic_bytes_per_particle_step": bpu2, "hbm_roofline_frac_step": (vo / world) * bpu2 / (peak * 1e9),
                          "sorts_total_incl_warmup": s2.sort_stats()[0]}
            s2.close()
            if rank == 0 and world == 1 and not args.no_cpu:
                others[wl]["cpu_baseline"] = cpu_baseline(wl, args.cpu_log2_particles, args.cpu_steps)

    if rank == 0:
        if args.workload == "gauss_fp":
            # bytes a step of S sweeps has to move as built: S fused passes of 24 / 32 / 40 B (first / middle / final) = 32*S.
            # SURVEY 8(d) counts 32*(S+1) (a separate k = 0 deposit pass, which is fused into the previous step's final pass
            # here): that figure is kept beside it, it can exceed 1 because those 32 B are never moved.
            bpu = 32.0 * mean_sweeps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "particles_per_gpu": per_gpu, "l2": "inputs_exceed_l2 (8*particles*arrays bytes >> 126 MB)",
                       "parallelism": f"particle-sharded x{world}, rho summed once per sweep: " + (
                           "NVLink peer-memory loads fused into the solve kernel (pg_peer.cuh)" if (world > 1 and sim_peer) else
                           "ncclAllReduce" if world > 1 else "single GPU")},
            "mean_sweeps_per_step": mean_sweeps, "particle_sweeps_per_s": value * mean_sweeps,
            "algorithmic_bytes_per_particle_step": bpu,
            "hbm_roofline_frac_step": (value / world) * bpu / (peak * 1e9),
            "hbm_roofline_frac_step_survey_8d_denominator": ((value / world) * 32.0 * (mean_sweeps + 1) / (peak * 1e9)) if args.workload == "gauss_fp" else None,
            "roofline": roofline, "fp64_pipe": fp64, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "sorts_in_timed_region": int(sorts_timed), "sorts_fused_into_the_passes_since_start": int(fused_main), "warm_regime": warm, "other_workloads": others,
        }
        print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, pg, sim, torch, barrier, max_over_ranks, stream, P, world):
    """Per step: the whole particle state of the shard goes host -> device and the stepped state comes back device -> host,
    through the C ABI.  Headline: picgolf_step_streamed (three device buffer sets, copies of neighbouring calls overlap the
    step; wall clock around Ke back-to-back calls and the final synchronize).  Beside it the serial
    picgolf_set_particles / picgolf_step(1) / picgolf_get_particles sequence of round 1."""
    n = sim.count
    ncomp = 5 if args.workload == "2d3v" else 2
    K = args.steps
    host = [torch.empty(n, dtype=torch.float64, pin_memory=True).numpy() for _ in range(ncomp)]
    outs = [[torch.empty(n, dtype=torch.float64, pin_memory=True).numpy() for _ in range(ncomp)] for _ in range(2)]
    got = sim.particles()
    for h, g in zip(host, got):
        h[:] = g
    del got
    Ke = max(1, args.e2e_steps)
    for i in range(2):  # builds the ring, touches every buffer
        sim.step_streamed(host, outs[i & 1])
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        sim.step_streamed(host, outs[i & 1])
    sim.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    barrier()
    ems = max_over_ranks(wall)
    e2e = {"value": P * Ke / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * ncomp * n, "d2h_bytes_per_step": 8 * ncomp * n,
           "steps": Ke, "ms_per_step": ems / Ke, "timer": "host wall clock around the calls and the final picgolf_synchronize (three CUDA streams: no single-stream event pair spans it)",
           "what": "per step: picgolf_step_streamed = upload of the whole shard state from pinned host memory, one PIC step, download of the stepped state; "
                   "uploads, steps and downloads of neighbouring calls overlap (3 device buffer sets)"}
    # the serial form: set -> step -> get, one after the other
    import ctypes as C
    lib, hnd = pg.load(), sim._h
    ptr = [h.ctypes.data_as(C.c_void_p) for h in outs[0]]
    Ks = max(1, min(3, Ke))

    def one():
        if ncomp == 2:
            pg._check(lib.picgolf_set_particles(hnd, host[0], host[1], n))
            pg._check(lib.picgolf_step(hnd, 1))
            pg._check(lib.picgolf_get_particles(hnd, ptr[0], ptr[1], n))
        else:
            pg._check(lib.picgolf_set_particles_2d3v(hnd, *host, n))
            pg._check(lib.picgolf_step(hnd, 1))
            pg._check(lib.picgolf_get_particles_2d3v(hnd, *ptr, n))
    one()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(Ks):
        one()
    f1.record(stream)
    barrier()
    sms = max_over_ranks(f0.elapsed_time(f1))
    e2e["serial_set_step_get"] = {"value": P * Ks / (sms * 1e-3), "unit": UNIT, "ms_per_step": sms / Ks, "steps": Ks,
                                  "what": "per step: picgolf_set_particles, picgolf_step(1), picgolf_get_particles, one after the other (CUDA events on the library's stream)"}
    del host, outs
    return e2e


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop, timed on the host cores
# ------------------------------------------------------------------------------------------------
def julia_probe():
    """Is the reference's own runtime on this box?  (probe before assuming.)"""
    exe = shutil.which("julia")
    ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref"))
    if exe:
        try:
            ver = subprocess.run([exe, "--version"], capture_output=True, text=True, timeout=60).stdout.strip()
        except Exception as e:  # noqa: BLE001
            ver = f"not runnable: {e}"
        return {"julia": exe, "version": ver, "baseline/_ref": ref,
                "note": "julia is present, but the reference scripts (/root/reference) do not travel to this box and must not be copied into the repo; the C restatement is timed"}
    return {"julia": None, "baseline/_ref": ref, "note": "no julia executable on PATH and no baseline/_ref: the reference (19 Julia scripts, no package) cannot run here"}


def cpu_baseline(workload, log2p, steps, warmup=1):
    from oracle import oracle as o

    P = 1 << log2p
    rng = np.random.default_rng(1234)
    if workload == "gauss_fp":
        N = 4096
        x0 = rng.random(P)
        v0 = np.where(np.arange(1, P + 1) > P / 2, 1.0, -1.0)
        fp = o.FixedPoint(x0, v0, N, 1 / (6 * N), 400.0, hw=6, rtol=1e-8)
        for _ in range(warmup):
            fp.step()
        t0 = time.perf_counter()
        sw = [fp.step()[2] for _ in range(steps)]
        dt = time.perf_counter() - t0
        cores, extra = 1, {"mean_sweeps_per_step": float(np.mean(sw))}
        sample = f"N=4096 P=2^{log2p} seeded-uniform two-stream, {steps} steps after {warmup} warm-up, single thread (the 1D scripts have no threading: JULIA_NUM_THREADS=1 is how the reference runs them)"
    elif workload == "ngp":
        N = 4096
        x = rng.random(P)
        v = np.where(np.arange(1, P + 1) > P / 2, 1.0, -1.0)
        for _ in range(warmup):
            o.ngp_step(x, v, N, 1 / (4 * N), 256.0 / P * N)
        t0 = time.perf_counter()
        for _ in range(steps):
            o.ngp_step(x, v, N, 1 / (4 * N), 256.0 / P * N)
        dt = time.perf_counter() - t0
        cores, extra = 1, {}
        sample = f"N=4096 P=2^{log2p} seeded-uniform two-stream, {steps} steps, single thread (NGPFourier.jl has no threading)"
    else:
        NX = NY = 256
        NG = math.sqrt(NX ** 2 + NY ** 2)
        n0 = 4 * math.pi ** 2
        vth = math.sqrt(n0) / NG
        st = [1 - rng.random(P), 1 - rng.random(P)] + [rng.standard_normal(P) * vth / math.sqrt(2) for _ in range(3)]
        Ex, Ey = np.zeros(NX * NY), np.zeros(NX * NY)
        cores = o.max_threads()
        args = (NX, NY, 1 / NG / (6 * vth), math.sqrt(n0) / 4, n0 / P * NX * NY)
        for _ in range(warmup):
            o.step_2d3v(*st, *args, Ex, Ey, nthreads=cores)
        t0 = time.perf_counter()
        for _ in range(steps):
            o.step_2d3v(*st, *args, Ex, Ey, nthreads=cores)
        dt = time.perf_counter() - t0
        extra = {}
        sample = f"256x256 P=2^{log2p}, {steps} steps, {cores} threads with per-thread grids (Electrostatic2D3V.jl:114,126-141: Threads.nthreads() = all host cores)"
    out = {"value": P * steps / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "sample_ms_per_step": dt / steps * 1e3,
           "sample_particles": P, "sample_steps": steps,
           "note": "C restatement of the Julia loop (oracle/picgolf_oracle.c, gcc -O2), not Julia", "reference_runtime_probe": julia_probe()}
    out.update(extra)
    return out


def run_reference(args):
    """The reference arm: the CPU implementation of the same path on this box's host cores.  It runs EXACTLY the requested
    `--warmup` + `--steps` steps, on a bounded sample of the workload: the particle count is cut (calibrated on one step so that
    the whole run takes ~30 s), every other parameter is the GPU arm's.  particle-steps/s is size-independent for these O(P)
    loops (the grid work is < 1 % at >= 16 particles per cell), and the line says what was run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, Wm = max(1, args.steps), max(0, args.warmup)
    cal = cpu_baseline(args.workload, 16, 1, warmup=1)  # calibration: seconds per particle-step on this box
    per = 1.0 / cal["value"]
    log2p = int(max(14, min(22, math.floor(math.log2(30.0 / ((K + Wm) * per))))))
    cpu = cpu_baseline(args.workload, log2p, K, warmup=Wm)
    per_gpu = 1 << args.log2_particles_per_gpu
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": Wm,
            "ms_per_step": cpu["sample_ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: bounded sample of the GPU arm's configuration -- 2^{log2p} particles instead of {per_gpu * max(world, args.gpus)}, "
                                   f"same grid, time step, stencil and tolerance; {K} timed steps after {Wm} warm-up steps on the host cores; {cpu['sample']}",
                       "sample_particles": cpu["sample_particles"], "sample_steps": K},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200, help="timed steps (default: ~1 s of device time at 2^28 particles)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="gauss_fp", choices=["gauss_fp", "ngp", "2d3v"])
    ap.add_argument("--log2-particles-per-gpu", type=int, default=28)
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--deposit-mode", default="auto", choices=["auto", "atomic", "sorted", "poly"],
                    help="gauss_fp only: force a deposit path (A/B runs); auto = what a caller gets")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", dest="others", action="store_false", help="skip the brief NGP and 2D3V runs")
    ap.add_argument("--no-warm", dest="warm", action="store_false", help="skip the warm-beam (saturated regime) run of config 4")
    ap.add_argument("--cpu-log2-particles", type=int, default=20)
    ap.add_argument("--cpu-steps", type=int, default=3)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
