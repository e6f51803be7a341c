#!/usr/bin/env python3
"""src/NGPFourier.jl (+ the K trace of NGPFourierWithDiagnostics.jl) on the GPU: N=128; P=64N; dt=1/4N; NT=1024; W=200."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particleincellcodegolf.jl_b200 as pg  # noqa: E402

sim = pg.ngp_fourier()                                         # line 1
P, NT = sim.cfg.P, sim.cfg.T
x = np.random.default_rng(0).random(P)                         # line 2: x=rand(P) (Julia's stream cannot be reproduced: passed in)
v = np.where(np.arange(1, P + 1) > P / 2, 1.0, -1.0)           # v=collect(1:P.>P/2).*2 .-1
sim.set_particles(x, v)
sim.step(NT)                                                   # lines 4-7
x, v = sim.particles()
n, E = sim.fields()
K, _ = sim.diagnostics()                                       # field, kinetic, total energy, mean momentum per step
print(f"total energy first/last {K[0, 2]:.6f} {K[-1, 2]:.6f}, max |mean momentum| {np.abs(K[:, 3]).max():.2e}")
np.savez("NGPFourier.npz", x=x, v=v, n=n, E=E, K=K)
