#!/usr/bin/env python3
"""Small runs through this round's new kernels, for compute-sanitizer: fp_pass_poly with the re-sort fused in (counting pass, multi-block
scan, scattering final pass; default and deterministic), the warp-shared deposit, solve1d_stock_kernel (N = 512, 2048, 8192), and the
leapfrog passes with the deposit carried across calls (TMA-staged and plain)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import particleincellcodegolf.jl_b200 as pg

part = sys.argv[1] if len(sys.argv) > 1 else "all"
for det in ((0, 1) if part in ("all", "poly") else ()):
    sim = pg.gaussian_fixed_point(N=512, P=(1 << 18) + 77, T=8, W=400.0, deposit_mode=pg.DEPOSIT_POLY, sort_every=1, deterministic=det)
    sim.init_synthetic(seed=3, vth=0.3)
    sim.step(5)
    x, v = sim.particles()
    assert np.isfinite(x).all() and sim.fused_sorts == 3
    sim.close()
for N in ((2048, 8192) if part in ("all", "solve") else ()):
    sim = pg.ngp_fourier(N=N, P=1 << 16, NT=8)
    sim.init_synthetic(seed=4)
    sim.step(2); sim.step(1); sim.step(2)
    sim.close()
if part in ("all", "tma"):
    sim = pg.ngp_fourier(N=1024, P=(1 << 20) + 5, NT=8)  # TMA-staged pass
    sim.init_synthetic(seed=5)
    sim.step(2); sim.step(1)
    x, v = sim.particles()
    assert np.isfinite(x).all()
    sim.close()
print("sanitize_fused done")
