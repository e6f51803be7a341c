import sys, time; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import particleincellcodegolf.jl_b200 as pg
for name, mk, init in (("C3 quiet N=64 P=2048 T=8192", lambda: pg.gaussian_fixed_point_quiet(), "quiet"),
                       ("C2 N=128 P=4096 T=1024", lambda: pg.gaussian_fixed_point(), "syn"),
                       ("C1 NGP N=128 P=8192 NT=1024", lambda: pg.ngp_fourier(), "syn")):
    sim = mk()
    (sim.init_quiet if init == "quiet" else sim.init_synthetic)()
    sim.step(8); sim.synchronize()
    T = sim.cfg.T - 8
    t0 = time.perf_counter(); sim.step(T); sim.synchronize(); dt = time.perf_counter() - t0
    D, sw = sim.diagnostics()
    print(f"{name}: {T} steps in {dt:.3f} s = {dt/T*1e6:.1f} us/step, {sim.cfg.P*T/dt/1e6:.1f} M particle-steps/s, mean sweeps {sw.mean():.2f}, launches {sim.launches}")
