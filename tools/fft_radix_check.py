#!/usr/bin/env python3
"""Dump solve outputs (1D for N = 16..8192, 2D 64x128) for a fixed random charge density; run once per library build
(PICGOLF_LIB=...) and compare the dumps: the radix-2^2 passes of pg_fft.cuh must reproduce the radix-2 ones bit for bit.

    python tools/fft_radix_check.py out.npz          # dump
    python tools/fft_radix_check.py a.npz b.npz      # compare
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    if len(sys.argv) == 3:
        a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
        ok = all(np.array_equal(a[k], b[k]) for k in a.files if not k.startswith("t_"))
        print("bit-identical" if ok else "DIFFERENT", {k: (float(a[k]), float(b[k])) for k in a.files if k.startswith("t_")})
        sys.exit(0 if ok else 1)
    import particleincellcodegolf.jl_b200 as pg
    rng = np.random.default_rng(0)
    out = {}
    for N in (16, 32, 64, 128, 1024, 2048, 4096, 8192):
        rho = 400.0 + rng.standard_normal(N)
        out[f"E{N}"] = pg.solve1d(rho)
    rho2 = 39.0 + rng.standard_normal((64, 128))
    ex, ey = pg.solve2d(rho2, 64, 128)
    out["Ex"], out["Ey"] = ex, ey
    # timing of the N = 4096 solve inside an NGP step loop (1 solve per step, tiny particle count)
    sim = pg.ngp_fourier(N=4096, P=4096, NT=8, W=256.0)
    sim.init_synthetic(seed=1)
    sim.step(50); sim.synchronize()
    t0 = time.perf_counter(); sim.step(2000); sim.synchronize()
    out["t_us_per_ngp_step_N4096_P4096"] = (time.perf_counter() - t0) / 2000 * 1e6
    np.savez(sys.argv[1], **out)
    print("wrote", sys.argv[1], out["t_us_per_ngp_step_N4096_P4096"])


if __name__ == "__main__":
    main()
