#!/usr/bin/env python3
"""Phase breakdown of solve1d_kernel from SM clock stamps (measurement build: tools/build_variants.sh solveprof "-DPG_SOLVE_PROF",
PICGOLF_LIB=particleincellcodegolf.jl_b200/lib/variants/libpicgolf_solveprof.so python tools/solve_prof.py)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particleincellcodegolf.jl_b200 as pg

lib = pg.load()
names = ["entry->rho loaded", "forward FFT", "divide by ik", "inverse FFT", "E out + norms", "block sums"]
for N in (128, 512, 1024, 2048, 4096, 8192):
    sim = pg.ngp_fourier(N=N, P=1 << 20, NT=64)
    sim.init_synthetic(seed=1)
    sim.step(8)
    sim.synchronize()
    out = (C.c_longlong * 8)()
    assert lib.picgolf_debug_solve_prof(out) == 0
    t = np.array(out[:7], dtype=np.int64)
    d = np.diff(t)
    print(f"N={N}: total {t[6] - t[0]} cycles = {(t[6] - t[0]) / 1.965e3:.1f} us at 1965 MHz; " + ", ".join(f"{n} {v}" for n, v in zip(names, d)))
    sim.close()
