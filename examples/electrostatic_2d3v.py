#!/usr/bin/env python3
"""src/Electrostatic2D3V.jl on the GPU: NX=NY=128; P=NX*NY*2^5; n0=4pi^2; vth=sqrt(n0)/NG; dt=1/NG/6vth; B0=sqrt(n0)/4; NS=2.
T is shortened from 2^15 steps to --steps (default 2^11) so that the example finishes in seconds."""
import argparse
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particleincellcodegolf.jl_b200 as pg  # noqa: E402
from particleincellcodegolf.jl_b200 import pic2d3v  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2 ** 11)
args = ap.parse_args()
NS = 2
TO = args.steps // NS                                           # snapshots kept (a power of two for the omega-k maps)
sim = pg.electrostatic_2d3v(T=TO, NS=NS, field_history=1)       # lines 23-25
P, vth = sim.cfg.P, sim.vth
rng = np.random.default_rng(0)                                  # lines 45-55 (Random.seed!(0); rand; erfinv): drawn here, passed in
x, y = 1 - rng.random(P), 1 - rng.random(P)                     # (0, 1]
vs = [rng.standard_normal(P) for _ in range(3)]
vs = [(a - a.mean()) / a.std() * vth / math.sqrt(2) for a in vs]
sim.set_particles(x, vs[0], y=y, vy=vs[1], vz=vs[2])
sim.step(args.steps)                                            # lines 120-176
K, _ = sim.diagnostics()                                        # K[ti,1:5]
Exs, Eys, phis = sim.snapshots("Exs"), sim.snapshots("Eys"), sim.snapshots("phis")   # lines 171-173, kept on the device
# lines 219-233: Z = log10.(sum(i->abs.(fft(F[:, i, :])), 1:size(F, 2))) and its y counterpart, transforms on the GPU
maps = {}
for name, F in (("Ex", Exs), ("Ey", Eys), ("phi", phis)):
    maps[name + "_kx"] = np.log10(np.maximum(pic2d3v.wk_spectrum(F, axis=0, mode=0), 1e-300))
    maps[name + "_ky"] = np.log10(np.maximum(pic2d3v.wk_spectrum(F, axis=1, mode=0), 1e-300))
print(f"rows {K.shape[0]}, total energy first/last {K[0, 2]:.6e} {K[-1, 2]:.6e}, mean momentum {K[-1, 3]:.2e} {K[-1, 4]:.2e}")
np.savez("Electrostatic2D3V.npz", K=K, Exs=Exs[:, :, -1], Eys=Eys[:, :, -1], phis=phis[:, :, -1], **maps)
