#!/usr/bin/env python3
"""src/2D3V.jl with the electrostatic field of src/PIC2D3V.jl (its commented-out lines 78-87) on the GPU: electrons + ions (M = 32),
BSplineWeighting{2}, Halton starts, ntskip = 4, ngskip = 2.  `--as-written` keeps update! as the reference has it (Exy accumulates
the solved fields, src/PIC2D3V.jl:294-297); the default here stores the field of the current step, which is what conserves energy."""
import argparse
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from particleincellcodegolf.jl_b200 import pic2d3v as p  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--NT", type=int, default=2 ** 10)
ap.add_argument("--NX", type=int, default=128)
ap.add_argument("--as-written", action="store_true")
args = ap.parse_args()
NX = NY = args.NX
Lx = Ly = 1.0
P = NX * NY * 16                                                # src/2D3V.jl:70
n0 = 4 * math.pi ** 2                                           # :78
dl = min(Lx / NX, Ly / NY)
vth = dl * math.sqrt(n0)                                        # :80 (debyeoverresolution = 1)
B0 = math.sqrt(n0) / 4                                          # :81
ntskip, M = 4, 32                                               # :84, :95
dt = dl / (6 * vth)                                             # :85
field = p.ElectrostaticField(NX, NY, Lx, Ly, dt=dt, B0x=B0, accumulate=args.as_written)       # :86
diagnostics = p.ElectrostaticDiagnostics(NX, NY, args.NT, ntskip, 2)                            # :87
shape = p.BSplineWeighting(2)                                                                   # :111
electrons = p.Species(P, vth, n0, shape, Lx=Lx, Ly=Ly, charge=-1, mass=1)                       # :114
ions = p.Species(P, vth / math.sqrt(M), n0, shape, Lx=Lx, Ly=Ly, charge=1, mass=M)              # :116
sim = p.Simulation([electrons, ions], field, diagnostics)       # Halton starts of Species(...) generated on the device
sim.loop(args.NT)                                               # :123-126  loop! + diagnose!
d = sim.scalars()
tot = d["kineticenergy"] + d["fieldenergy"]
print(f"rows {len(tot)}, (field+kinetic)/initial at the end {tot[-1] / tot[0]:.4f}, |momentum|/characteristic "
      f"{np.abs(d['particlemomentum'][-1]).max() / d['characteristicmomentum'][0].max():.2e}, sorts/misses {sim.sort_stats()}")
out = dict(kineticenergy=d["kineticenergy"], fieldenergy=d["fieldenergy"], particlemomentum=d["particlemomentum"],
           characteristicmomentum=d["characteristicmomentum"])
for name in ("Exs", "Eys", "phis"):                             # plotfields :1483-1493: abs.(fft(F))[2:kxind, 1, 1:wind] and its y counterpart
    out[name + "_last"] = sim.history(name)[:, :, -1]
    out[name + "_wkx"] = np.log10(np.maximum(sim.spectrum(name, axis=0, mode=1), 1e-300))
    out[name + "_wky"] = np.log10(np.maximum(sim.spectrum(name, axis=1, mode=1), 1e-300))
np.savez("PIC2D3V_electrostatic.npz", **out)
