#!/usr/bin/env python3
"""bench.py -- particle-steps/s of the per-timestep PIC loop (deposit + solve + push) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload gauss_fp|ngp|2d3v] [--impl reference]

Headline workload (BASELINE.json configs[3]): scaled synthetic Gaussian fixed-point two-stream 1D1V,
N=4096 cells, 2^28 particles per GPU (weak scaling), src/GaussianFixedPointQuiet.jl loop body with the
+-6 stencil and l=1e-8, seeded-uniform two-stream start (the GaussianFixedPoint.jl x=rand(P) pattern).
A "step" is one PIC time step = S fixed-point sweeps (S reported).  One JSON line on stdout.

Timing: W untimed warm-up steps, then exactly K steps bracketed by barrier + synchronize; CUDA events
on the library's own stream; max over ranks.  Inputs (8 GiB of particle state per GPU) exceed L2.
`value` has the particle state resident in HBM; `e2e` moves the whole particle state host->device and
device->host through the C ABI every step (pinned host buffers).
"""
import argparse
import json
import math
import os
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
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [s.strip() for s in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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


def init_sim(sim, workload, seed=1234):
    sim.init_synthetic(seed=seed, vth=getattr(sim, "vth", 0.0))


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
    torch.cuda.set_device(local)
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

    # ---- device-resident timing -------------------------------------------------------------
    sim.step(Wm)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    sorts0 = sim.sort_stats()[0]
    l0 = sim.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.step(K)  # one C-ABI call enqueues all K steps (no host sync inside)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = sim.launches - l0
    sorts_timed = sim.sort_stats()[0] - sorts0
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    P = sim.cfg.P
    value = P * K / (ms * 1e-3)
    D, sw = sim.diagnostics()
    sweeps = sw[Wm:Wm + K].astype(float) if args.workload == "gauss_fp" else np.ones(K)
    mean_sweeps = float(sweeps.mean())

    # ---- roofline of the dominant kernel (the particle pass), timed live with CUDA events ------
    sim.stage_timing(True)
    sim.stage_times(reset=True)
    nroof = 40 if (args.workload == "gauss_fp" and sim.deposit_path == pg.DEPOSIT_POLY) else 12  # long enough to contain a re-sort
    s0 = sim.sort_stats()[0]
    sim.step(nroof)
    st = sim.stage_times(reset=True)
    sim.stage_timing(False)
    _, sw2 = sim.diagnostics()
    if args.workload == "gauss_fp":
        passes = float(sw2[Wm + K:Wm + K + nroof].sum())  # S solves -> S particle passes per step (the k=0 pass is fused into the previous step's final pass)
        alg_bytes_launch = 32.0 * per_gpu
        poly = sim.deposit_path == pg.DEPOSIT_POLY
        kernel = ("fp_pass_poly<FIRST> (per-cell gather polynomial + implicit-midpoint update + register moment deposit)" if poly else
                  "fp_pass_sorted<FIRST,NP=2> (gather + implicit-midpoint update + deposit; 3 variants: first/middle/final pass)")
    elif args.workload == "ngp":
        passes = float(nroof + 1)
        alg_bytes_launch = 32.0 * per_gpu
        kernel = "lf_pass_ngp_tma (TMA-staged tiles: drift + kick + drift + NGP deposit)"
    else:
        passes = float(nroof)
        alg_bytes_launch = 80.0 * per_gpu
        kernel = "particles_2d3v_tiled (gather + boris + move + CIC deposit)"
    launch_ms = st["particles"] / passes
    peak, peak_src = peaks()
    achieved = alg_bytes_launch / (launch_ms * 1e-3) / 1e9
    traffic = None
    try:  # per-launch DRAM bytes of the same kernel from the committed ncu --set full capture (only valid at 2^28/GPU)
        key = "gauss_fp_poly" if args.workload == "gauss_fp" and sim.deposit_path == pg.DEPOSIT_POLY else args.workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[key]
        if per_gpu == 1 << 28:
            traffic = tj["traffic_bytes"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": traffic, "launch_ms": launch_ms, "algorithmic_bytes_per_launch": alg_bytes_launch,
                "stage_ms_per_step": {k: v / nroof for k, v in st.items()},
                "resorts_in_stage_window": sim.sort_stats()[0] - s0, "stage_window_steps": nroof}

    # ---- FP64 pipe view of the same kernel (the binding limit of the erf-shape passes) ----------
    fp64 = None
    if args.workload == "gauss_fp":
        peak_tf = pg.fp64_peak_tflops()
        # FP64 instructions per particle-pass measured with ncu (profiles/): 255 for a fp_pass_sorted pass with two
        # stencils (middle passes and the final pass with the fused first deposit of the next step); 68 for fp_pass_poly
        # (FP64 pipe active 57.6 % of 3.25 M cycles = 132 warp-instructions per row of 64 particles in the first and
        # middle passes, 144 in the final pass: mean over a 3-sweep step 68 per particle)
        per_pass = 68 if sim.deposit_path == pg.DEPOSIT_POLY else 255
        fp64_inst = float(sum(per_pass * int(s_) for s_ in sw2[Wm + K:Wm + K + nroof])) * per_gpu
        ach = fp64_inst * 2.0 / (st["particles"] * 1e-3) / 1e12  # counted as 2 flops per lane-instruction (FMA-equivalent)
        fp64 = {"measured_peak_tflops": peak_tf, "achieved_tflops_fma_equiv": ach, "frac": ach / peak_tf,
                "fp64_instructions_per_particle_pass": per_pass,
                "note": "FP64 lane-instructions of the pass kernels x2 / kernel time (the second bound of the erf-shape passes besides HBM)"}

    # ---- end to end through the C ABI with host buffers ------------------------------------------
    e2e = None
    if not args.no_e2e:
        n = sim.count
        ncomp = 5 if args.workload == "2d3v" else 2
        host = [torch.empty(n, dtype=torch.float64, pin_memory=True).numpy() for _ in range(ncomp)]
        got = sim.particles()
        for h, g in zip(host, got):
            h[:] = g
        del got
        lib, hnd = pg.load(), sim._h
        import ctypes as C
        ptr = [h.ctypes.data_as(C.c_void_p) for h in host]
        Ke = max(1, min(K, args.e2e_steps))

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
        t0 = time.perf_counter()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(Ke):
            one()
        f1.record(stream)
        barrier()
        ems = f0.elapsed_time(f1)  # device time between the first H2D and the last D2H on the library's stream
        wall = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([ems], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": P * Ke / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * ncomp * n, "d2h_bytes_per_step": 8 * ncomp * n,
               "steps": Ke, "ms_per_step": ems / Ke, "wall_ms_per_step": wall / Ke,
               "what": "per step: picgolf_set_particles (pinned host -> HBM), picgolf_step(1), picgolf_get_particles (HBM -> host)"}
        # second reading of "inputs in, result out": the particle state goes in every step, only the step's fields
        # (rho, E) and its diagnostics row come back
        def one_fields():
            if ncomp == 2:
                pg._check(lib.picgolf_set_particles(hnd, host[0], host[1], n))
            else:
                pg._check(lib.picgolf_set_particles_2d3v(hnd, *host, n))
            pg._check(lib.picgolf_step(hnd, 1))
            fl = sim.fields()
            sim.diagnostics()
            return sum(a.nbytes for a in fl) + 40
        one_fields()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(Ke):
            d2h = one_fields()
        f1.record(stream)
        barrier()
        fms = f0.elapsed_time(f1)
        if world > 1:
            t = torch.tensor([fms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fms = float(t.item())
        e2e["state_in_fields_out"] = {"value": P * Ke / (fms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * ncomp * n,
                                      "d2h_bytes_per_step": int(d2h), "ms_per_step": fms / Ke,
                                      "what": "per step: picgolf_set_particles, picgolf_step(1), picgolf_get_fields + picgolf_get_diagnostics"}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) -------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args.workload, args.cpu_log2_particles, args.cpu_steps)

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
            mso = g0.elapsed_time(g1)
            if world > 1:
                t = torch.tensor([mso], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                mso = float(t.item())
            vo = s2.cfg.P * Ko / (mso * 1e-3)
            others[wl] = {"workload": name2, "value": vo, "unit": UNIT, "steps": Ko, "ms_per_step": mso / Ko,
                          "algorithmic_bytes_per_particle_step": bpu2, "hbm_roofline_frac_step": (vo / world) * bpu2 / (peak * 1e9),
                          "sorts_total_incl_warmup": s2.sort_stats()[0]}
            s2.close()

    if rank == 0:
        if args.workload == "gauss_fp":
            bpu = 32.0 * (mean_sweeps + 1)
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
            "roofline": roofline, "fp64_pipe": fp64, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "sorts_in_timed_region": int(sorts_timed), "other_workloads": others,
        }
        print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop, timed on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_baseline(workload, log2p, steps):
    from oracle import oracle as o

    P = 1 << log2p
    rng = np.random.default_rng(1234)
    if workload == "gauss_fp":
        N = 4096
        x0 = rng.random(P)
        v0 = np.where(np.arange(1, P + 1) > P / 2, 1.0, -1.0)
        fp = o.FixedPoint(x0, v0, N, 1 / (6 * N), 400.0, hw=6, rtol=1e-8)
        fp.step()
        t0 = time.perf_counter()
        sw = [fp.step()[2] for _ in range(steps)]
        dt = time.perf_counter() - t0
        cores, extra = 1, {"mean_sweeps_per_step": float(np.mean(sw))}
        sample = f"N=4096 P=2^{log2p} seeded-uniform two-stream, {steps} steps after 1 warm-up, single thread (the 1D scripts have no threading)"
    elif workload == "ngp":
        N = 4096
        x = rng.random(P)
        v = np.where(np.arange(1, P + 1) > P / 2, 1.0, -1.0)
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
        o.step_2d3v(*st, *args, Ex, Ey, nthreads=cores)
        t0 = time.perf_counter()
        for _ in range(steps):
            o.step_2d3v(*st, *args, Ex, Ey, nthreads=cores)
        dt = time.perf_counter() - t0
        extra = {}
        sample = f"256x256 P=2^{log2p}, {steps} steps, {cores} threads with per-thread grids (Electrostatic2D3V.jl:114,126-141)"
    out = {"value": P * steps / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "sample_ms_per_step": dt / steps * 1e3,
           "note": "C restatement of the Julia loop (oracle/picgolf_oracle.c, gcc -O2), not Julia: julia is not installed"}
    out.update(extra)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, Wm = args.steps, max(args.warmup, 1)
    cpu = cpu_baseline(args.workload, args.cpu_log2_particles, max(1, min(K, args.cpu_steps)))
    per_gpu = 1 << args.log2_particles_per_gpu
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": Wm,
            "ms_per_step": cpu["sample_ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: same configuration as the GPU arm ({per_gpu * max(world, args.gpus)} particles), timed on a bounded sample",
                       "sample": cpu["sample"]},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="gauss_fp", choices=["gauss_fp", "ngp", "2d3v"])
    ap.add_argument("--log2-particles-per-gpu", type=int, default=28)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--deposit-mode", default="auto", choices=["auto", "atomic", "sorted", "poly"],
                    help="gauss_fp only: force a deposit path (A/B runs); auto = what a caller gets")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", dest="others", action="store_false", help="skip the brief NGP and 2D3V runs")
    ap.add_argument("--cpu-log2-particles", type=int, default=20)
    ap.add_argument("--cpu-steps", type=int, default=3)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
