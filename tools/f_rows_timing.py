#!/usr/bin/env python3
"""Device-resident throughput of the SURVEY 8(f) rank 1-2 schemes, which run on the any-order kernels only:
Simpson-1/3 fixed point with the Gaussian and the area shape (src/GaussianFixedPointQuietSimpson13.jl,
src/AreaFixedPointQuietSimpson13.jl) and the magnetised 1D2V codes (src/NGP1D2V.jl, src/NGP1D2V2S.jl), scaled to
2^log2p particles on an N-cell grid.  CUDA events on the handle's stream; particle arrays far larger than L2.
Run on a B200:  python tools/f_rows_timing.py [--log2p 24] [--grid 4096] [--steps 8]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2p", type=int, default=24)
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--schemes", default="simpson_gauss,simpson_area,1d2v,1d2v2s")
    args = ap.parse_args()
    import torch

    import particleincellcodegolf.jl_b200 as pg
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    P, N = 1 << args.log2p, args.grid
    rng = np.random.default_rng(5)
    for scheme in args.schemes.split(","):
        T = args.steps + 8
        if scheme == "simpson_gauss":
            sim = pg.gaussian_fixed_point_quiet_simpson13(N=N, P=P, T=T, l=1e-8)
            sim.init_synthetic(seed=11)
            nparts, bytes_per = P, None
        elif scheme == "simpson_area":
            sim = pg.area_fixed_point_quiet_simpson13(N=N, P=P, T=T, l=1e-8)
            sim.init_synthetic(seed=11)
            nparts, bytes_per = P, None
        elif scheme == "1d2v":
            sim = pg.ngp_1d2v(N=N, P=P, T=T * 16, TO=T)
            sim.set_particles(rng.random(P), sim.vth * rng.standard_normal(P), vy=sim.vth * rng.standard_normal(P))
            nparts, bytes_per = P, 48.0  # x, vx, vy read + written
        elif scheme == "1d2v2s":
            sim = pg.ngp_1d2v_2s(N=N, P=P // 2, T=T * 32, TO=T)
            sim.set_particles(rng.random(P), sim.vth * rng.standard_normal(P), vy=sim.vth * rng.standard_normal(P))
            nparts, bytes_per = P, 48.0
        else:
            raise SystemExit("unknown scheme " + scheme)
        stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", 0))
        sim.step(3)
        sim.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = sim.launches
        e0.record(stream)
        sim.step(args.steps)
        e1.record(stream)
        sim.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        l1 = sim.launches
        sim.stage_timing(True)
        sim.stage_times(reset=True)
        sim.step(4)
        st = sim.stage_times(reset=True)
        sim.stage_timing(False)
        row = {"scheme": scheme, "N": N, "particles": nparts, "ms_per_step": ms, "particle_steps_per_s": nparts / (ms * 1e-3),
               "launches_per_step": (l1 - l0) / args.steps, "stage_ms_per_step": {k: v / 4 for k, v in st.items()}}
        if bytes_per:
            row["hbm_frac_at_%dB" % bytes_per] = nparts * bytes_per / (ms * 1e-3) / (hbm * 1e9)
        if scheme.startswith("simpson"):
            _, sw = sim.diagnostics()
            row["sweeps_last_steps"] = [int(v) for v in sw[:args.steps + 7][-4:]]
        print(json.dumps(row), flush=True)
        sim.close()


if __name__ == "__main__":
    main()
