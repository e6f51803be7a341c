#!/usr/bin/env python3
"""src/GaussianFixedPointQuiet.jl on the GPU: N=64; P=32N; dt=1/6N; T=2^13; W=32pi^2/3; l=4eps(); stencil -7:7; quiet start."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particleincellcodegolf.jl_b200 as pg  # noqa: E402

sim = pg.gaussian_fixed_point_quiet()          # lines 1-6
sim.init_quiet()                               # lines 2-3: x=(bitreverse.(0:P-1).+2.0^63)/2.0^64; v=+-1 by halves
T, dt, W = sim.cfg.T, sim.cfg.dt, sim.cfg.W
sim.step(T)                                    # lines 8-15 (no host synchronisation inside)
D, sweeps = sim.diagnostics()                  # D[t,1:4] as line 11/13 forms it

t = np.arange(1, T + 1) * dt                   # lines 16-20: la(x)=log10(abs(x)); the analytic line through index T/8
x = 2 * math.pi / math.sqrt(W / 2)
gamma = math.sqrt(-(x ** 2 + 1 - math.sqrt(4 * x ** 2 + 1))) * math.sqrt(W / 2) / math.log(10)
la = np.log10(np.abs(D[:, 0]))
sel = (t > 1) & (t < 8)
slope = np.polyfit(t[sel], la[sel], 1)[0]
print(f"field-energy growth: fitted {slope:.4f} decades/time, analytic 2*gamma = {2 * gamma:.4f}")
print(f"max |momentum| {np.abs(D[:, 3]).max():.2e}, max |1 - total energy| {np.abs(1 - D[:, 2]).max():.2e}, sweeps/step {np.bincount(sweeps)}")
np.savez("GaussianFixedPointQuiet.npz", t=t, D=D, sweeps=sweeps, slope=slope, two_gamma=2 * gamma)
