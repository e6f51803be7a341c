#!/usr/bin/env python3
"""Device-resident timing of the PIC2D3V.jl electrostatic path (include/picgolf_es.h): particle-steps/s of loop! + diagnose!
for each shape, CUDA events on the handle's stream, inputs far larger than L2.  Algorithmic bytes: 80 B per particle-step
(x, y, vx, vy, vz read + written).  Run on a B200:  python tools/es_timing.py [--log2p 25] [--grid 256]"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2p", type=int, default=25, help="particles per species (two species)")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--shapes", default="0,1,12,13,15")
    ap.add_argument("--sort-every", default="0,-1", help="0 = auto (tile-sorted at this size), -1 = any-order kernel, n = sort every n steps")
    args = ap.parse_args()
    import torch

    import particleincellcodegolf.jl_b200 as pg
    import particleincellcodegolf.jl_b200.pic2d3v as es
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    NX = NY = args.grid
    P = 1 << args.log2p
    n0 = 4 * math.pi ** 2
    dl = 1.0 / NX
    vth = dl * math.sqrt(n0)
    dt = dl / (6 * vth)
    out = []
    for sort_every in [int(v) for v in args.sort_every.split(",")]:
      for code in [int(c) for c in args.shapes.split(",")]:
          shape = es.NGPWeighting() if code == 0 else es.AreaWeighting() if code == 1 else es.BSplineWeighting(code - 10)
          plasma = [es.Species(P, vth, n0, shape, Lx=1.0, Ly=1.0, charge=-1, mass=1), es.Species(P, vth / math.sqrt(32), n0, shape, Lx=1.0, Ly=1.0, charge=1, mass=32)]
          sim = es.Simulation(plasma, es.ElectrostaticField(NX, NY, 1.0, 1.0, dt=dt, B0x=math.sqrt(n0) / 4, accumulate=False),
                              es.ElectrostaticDiagnostics(NX, NY, 4 * (args.steps + 4), 4, 2, history=True), sort_every=sort_every)
          st = C_stream(sim)
          stream = torch.cuda.ExternalStream(st, device=torch.device("cuda", 0))
          sim.loop(3)
          sim.synchronize()
          e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          l0 = sim.launches
          e0.record(stream)
          sim.loop(args.steps)
          e1.record(stream)
          sim.synchronize()
          torch.cuda.synchronize()
          ms = e0.elapsed_time(e1) / args.steps
          rate = 2 * P / (ms * 1e-3)
          sc = sim.scalars()
          row = {"sort_every": sort_every, "sorts_slow": sim.sort_stats(), "shape": code, "grid": NX, "particles": 2 * P, "ms_per_step": ms, "particle_steps_per_s": rate,
                 "hbm_frac_at_80B": rate * 80 / (hbm * 1e9), "launches_per_step": (sim.launches - l0) / args.steps,
                 "energy_drift": float((sc["kineticenergy"][-1] + sc["fieldenergy"][-1]) / (sc["kineticenergy"][0] + sc["fieldenergy"][0]) - 1)}
          print(json.dumps(row), flush=True)
          out.append(row)
          sim.close()
    return out


def C_stream(sim):
    import ctypes as C
    v = C.c_void_p()
    rc = sim._lib.picgolf_es_get_stream(sim._h, C.byref(v))
    assert rc == 0
    return v.value


if __name__ == "__main__":
    main()
