#!/usr/bin/env python3
"""Config 4 (N=4096, 2^28 particles) in steady state for cold and warm beams: ms/step against the re-sort interval (0 = adaptive), with the
re-sort fused into the passes (default) and with the stand-alone counting sort (PICGOLF_FUSED_SORT=0).  40 steps of warm-up (the adaptive
interval settles), 24 timed.  usage: python tools/fused_sort_timing.py [log2P]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import particleincellcodegolf.jl_b200 as pg

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
N, P = 4096, 1 << lg
for fused in (1,):
    os.environ["PICGOLF_FUSED_SORT"] = str(fused)
    for vth in (0.0, 0.05, 0.3, 1.0):
        for se in ((0, 1) if fused else (0,)):
            sim = pg.gaussian_fixed_point(N=N, P=P, T=128, W=400.0, sort_every=se)
            sim.init_synthetic(seed=99, vth=vth)
            sim.step(40)
            sim.synchronize()
            st = torch.cuda.ExternalStream(sim.stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0, f0 = sim.sort_stats()
            K = 24
            e0.record(st); sim.step(K); e1.record(st); sim.synchronize(); torch.cuda.synchronize()
            s1, f1 = sim.sort_stats()
            sw = sim.diagnostics()[1][40:40 + K]
            print(f"fused={fused} vth={vth} sort_every={se}: {e0.elapsed_time(e1) / K:.3f} ms/step, sweeps {sw.mean():.2f}, sorts {s1 - s0} in {K} steps, "
                  f"flushes/particle-step {(f1 - f0) / (P * K):.2e}", flush=True)
            sim.close()
