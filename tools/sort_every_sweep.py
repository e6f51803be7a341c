import sys, os; sys.path.insert(0, os.getcwd())
import torch, numpy as np
import particleincellcodegolf.jl_b200 as pg
for se in (0, 16, 32):
    sim = pg.electrostatic_2d3v(NX=256, NY=256, P=1<<28, T=8, NS=1, sort_every=se)
    sim.init_synthetic(seed=1, vth=sim.vth)
    sim.step(4); sim.synchronize()
    st = torch.cuda.ExternalStream(sim.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 96
    e0.record(st); sim.step(K); e1.record(st); sim.synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/K
    sorts, slow = sim.sort_stats()
    print(f"sort_every {se}: {ms:.3f} ms/step, {sim.cfg.P/ms/1e6:.1f} G/s, sorts {sorts}, slow-path deposits {slow} ({slow/(sim.cfg.P*K):.2e} per particle-step)", flush=True)
    sim.close()
