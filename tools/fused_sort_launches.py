#!/usr/bin/env python3
"""A few steps of config 4 with a fused re-sort in every step, for an ncu launch list (per-kernel durations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particleincellcodegolf.jl_b200 as pg
vth = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
se = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sim = pg.gaussian_fixed_point(N=4096, P=1 << 28, T=32, W=400.0, sort_every=se)
sim.init_synthetic(seed=99, vth=vth)
sim.step(6)
sim.synchronize()
print(sim.sort_stats(), sim.fused_sorts)
