#!/usr/bin/env python3
"""Run under torchrun (one rank per GPU): the N-rank sharded run must reproduce the 1-rank run.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py

Checks, for the Gaussian fixed point (N=4096), NGP and 2D3V schemes: every rank's shard of x, v equals the
corresponding slice of a single-GPU run (rank 0 runs it too) within round-off, rho/E agree, sweep counts are equal
on all ranks, diagnostics are the global sums.  The charge grid is integer fixed point, so rho is expected to be
BIT-IDENTICAL between 1 and N GPUs in the order-free (atomic) mode."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particleincellcodegolf.jl_b200 as pg  # noqa: E402
from particleincellcodegolf.jl_b200 import distributed as pgd  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check_esfield(rank, world, local):
    """PIC2D3V.jl electrostatic path (include/picgolf_es.h): two species sharded over the ranks, device-side Halton start
    (global mean / variance all-reduced), rho all-reduced every step; against the same run on one GPU."""
    import math

    from particleincellcodegolf.jl_b200 import pic2d3v as es

    NX, NY, Lx, Ly, NT = 64, 32, 2.0, 1.0, 8
    P = (1 << 18) + 3  # not divisible by the rank count
    n0 = 4 * math.pi ** 2
    dl = min(Lx / NX, Ly / NY)
    vth = dl * math.sqrt(n0)

    def build(r, w):
        plasma = [es.Species(P, vth, n0, es.BSplineWeighting(2), Lx=Lx, Ly=Ly, charge=-1, mass=1),
                  es.Species(P, vth / 4, n0, es.AreaWeighting(), Lx=Lx, Ly=Ly, charge=1, mass=16)]
        sim = es.Simulation(plasma, es.ElectrostaticField(NX, NY, Lx, Ly, dt=dl / (6 * vth), B0x=math.sqrt(n0) / 4, B0z=0.5, accumulate=False),
                            es.ElectrostaticDiagnostics(NX, NY, NT, 2, 2), device=local, rank=r, nranks=w)
        if w > 1:
            sim.connect()
            sim.init_particles()
        return sim

    sim, ref = build(rank, world), build(0, 1)
    good = True
    for s in range(2):  # the start itself: same particles as the single-GPU start
        f, c = sim.ranges[s]
        for a, b in zip(sim.species(s), ref.species(s)):
            good = good and rel(a, b[f:f + c]) < 1e-12
    sim.loop(NT); ref.loop(NT)
    fa, fb = sim.fields(), ref.fields()
    e = dict(rho=rel(fa["rho"], fb["rho"]), Ex=rel(fa["Ex"], fb["Ex"]), Exy=rel(fa["Exy_y"], fb["Exy_y"]))
    for s in range(2):
        f, c = sim.ranges[s]
        for k, (a, b) in enumerate(zip(sim.species(s), ref.species(s))):
            e[f"s{s}c{k}"] = rel(a, b[f:f + c])
    sa, sb = sim.scalars(), ref.scalars()
    e["ke"] = rel(sa["kineticenergy"], sb["kineticenergy"]); e["fe"] = rel(sa["fieldenergy"], sb["fieldenergy"])
    e["cmom"] = rel(sa["characteristicmomentum"], sb["characteristicmomentum"])
    e["Exs"] = rel(sim.history("Exs"), ref.history("Exs"))
    good = good and max(e.values()) < 1e-10 and len(sa["kineticenergy"]) == NT // 2
    print(f"[rank {rank}] esfield {e} ok={good}", flush=True)
    sim.close(); ref.close()
    return good


def main():
    rank, world, local = pgd.env_rank()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    if os.environ.get("PICGOLF_CHECK_ONLY") == "es":
        ok = check_esfield(rank, world, local)
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
        if rank == 0:
            print("MULTIGPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
        sys.exit(0 if int(t.item()) == 1 else 1)
    peer_seen = None
    for mode in (pg.DEPOSIT_ATOMIC, pg.DEPOSIT_AUTO, pg.DEPOSIT_POLY):
        N, P, steps = 4096, 1 << 21, 6
        sim = pg.gaussian_fixed_point(N=N, P=P, T=16, W=400.0, rank=rank, nranks=world, device=local, deposit_mode=mode, sort_every=3)
        pgd.connect(sim)
        sim.init_synthetic(seed=5)
        sim.step(steps)
        x, v = sim.particles()
        rho, E = sim.fields()
        D, sw = sim.diagnostics()
        ref = pg.gaussian_fixed_point(N=N, P=P, T=16, W=400.0, device=local, deposit_mode=mode, sort_every=3)
        ref.init_synthetic(seed=5)
        ref.step(steps)
        xr, vr = ref.particles()
        rr, Er = ref.fields()
        Dr, swr = ref.diagnostics()
        f, c = sim.first, sim.count
        e = dict(x=rel(x, xr[f:f + c]), v=rel(v, vr[f:f + c]), rho=rel(rho, rr), E=rel(E, Er), D=rel(D[:, :3], Dr[:, :3]))
        good = max(e.values()) < 1e-10 and np.array_equal(sw, swr) and abs(D[:, 3] - Dr[:, 3]).max() < 1e-13
        if mode == pg.DEPOSIT_ATOMIC:
            good = good and np.array_equal(rho, rr)  # integer accumulation: independent of the GPU count
        peer_seen = sim.peer_status
        good = good and not peer_seen[1]
        print(f"[rank {rank}] fixed-point mode={mode} {e} sweeps={list(sw)} bit_equal_rho={np.array_equal(rho, rr)} peer(enabled,timed_out)={peer_seen} ok={good}", flush=True)
        ok &= good
        sim.close(); ref.close()
    # NGP
    sim = pg.ngp_fourier(N=4096, P=1 << 21, NT=8, W=256.0, rank=rank, nranks=world, device=local)
    pgd.connect(sim)
    sim.init_synthetic(seed=6)
    sim.step(4)
    x, v = sim.particles(); rho, E = sim.fields(); R = sim.raw_diagnostics()
    ref = pg.ngp_fourier(N=4096, P=1 << 21, NT=8, W=256.0, device=local)
    ref.init_synthetic(seed=6); ref.step(4)
    xr, vr = ref.particles(); rr, Er = ref.fields(); Rr = ref.raw_diagnostics()
    f, c = sim.first, sim.count
    good = np.array_equal(rho, rr) and np.array_equal(E, Er) and np.array_equal(x, xr[f:f + c]) and np.array_equal(v, vr[f:f + c]) \
        and rel(R[:, :3], Rr[:, :3]) < 1e-12
    print(f"[rank {rank}] ngp bit-identical={good}", flush=True)
    ok &= good
    sim.close(); ref.close()
    # 2D3V
    for mode in (pg.DEPOSIT_ATOMIC, pg.DEPOSIT_AUTO):
        sim = pg.electrostatic_2d3v(NX=128, NY=128, P=1 << 21, T=8, NS=1, rank=rank, nranks=world, device=local, deposit_mode=mode, sort_every=3)
        pgd.connect(sim)
        sim.init_synthetic(seed=7, vth=sim.vth)
        sim.step(5)
        got = sim.particles(); fld = sim.fields(); K, _ = sim.diagnostics()
        ref = pg.electrostatic_2d3v(NX=128, NY=128, P=1 << 21, T=8, NS=1, device=local, deposit_mode=mode, sort_every=3)
        ref.init_synthetic(seed=7, vth=ref.vth); ref.step(5)
        gr = ref.particles(); fr = ref.fields(); Kr, _ = ref.diagnostics()
        f, c = sim.first, sim.count
        e = dict(x=rel(got[0], gr[0][f:f + c]), vx=rel(got[2], gr[2][f:f + c]), rho=rel(fld[0], fr[0]), Ex=rel(fld[1], fr[1]), K=rel(K[:, :3], Kr[:, :3]))
        good = max(e.values()) < 1e-10
        print(f"[rank {rank}] 2d3v mode={mode} {e} ok={good}", flush=True)
        ok &= good
        sim.close(); ref.close()
    # "next" rows: Simpson-1/3 fixed point (3-row grid all-reduce) and the 1D2V Boris code
    for ctor in (pg.gaussian_fixed_point_quiet_simpson13, pg.area_fixed_point_quiet_simpson13):
        sim = ctor(N=256, P=1 << 16, T=8, W=400.0, l=1e-9, rank=rank, nranks=world, device=local)
        pgd.connect(sim)
        sim.init_synthetic(seed=8)
        sim.step(3)
        x, v = sim.particles(); rho, E = sim.fields(); D, sw = sim.diagnostics()
        ref = ctor(N=256, P=1 << 16, T=8, W=400.0, l=1e-9, device=local)
        ref.init_synthetic(seed=8); ref.step(3)
        xr, vr = ref.particles(); rr, Er = ref.fields(); Dr, swr = ref.diagnostics()
        f, c = sim.first, sim.count
        good = np.array_equal(x, xr[f:f + c]) and np.array_equal(v, vr[f:f + c]) and np.array_equal(rho, rr) and np.array_equal(E, Er) \
            and np.array_equal(sw, swr) and rel(D[:, :3], Dr[:, :3]) < 1e-12
        print(f"[rank {rank}] {ctor.__name__} bit-identical={good} sweeps={list(sw)}", flush=True)
        ok &= good
        sim.close(); ref.close()
    rng = np.random.default_rng(3)
    P = 1 << 16
    sim = pg.ngp_1d2v(N=512, P=P, T=64, TO=16, rank=rank, nranks=world, device=local)
    pgd.connect(sim)
    xg, vxg, vyg = rng.random(P), sim.vth * rng.standard_normal(P), sim.vth * rng.standard_normal(P)
    f, c = sim.first, sim.count
    sim.set_particles(xg[f:f + c], vxg[f:f + c], vy=vyg[f:f + c])
    sim.step(8)
    got = sim.particles(); fld = sim.fields(); D, _ = sim.diagnostics(); Es = sim.field_history()
    ref = pg.ngp_1d2v(N=512, P=P, T=64, TO=16, device=local)
    ref.set_particles(xg, vxg, vy=vyg); ref.step(8)
    gr = ref.particles(); fr = ref.fields(); Dr, _ = ref.diagnostics(); Esr = ref.field_history()
    good = all(np.array_equal(a, b[f:f + c]) for a, b in zip(got, gr)) and np.array_equal(fld[0], fr[0]) and np.array_equal(Es, Esr) \
        and rel(D, Dr) < 1e-12
    print(f"[rank {rank}] ngp_1d2v bit-identical={good}", flush=True)
    ok &= good
    sim.close(); ref.close()
    # two-species 1D2V: the species boundary falls inside rank 0's shard for world = 3.. and between shards for world = 2
    P2 = 1 << 15
    sim = pg.ngp_1d2v_2s(N=256, P=P2, T=64, TO=16, M=8.0, rank=rank, nranks=world, device=local)
    pgd.connect(sim)
    xg, vxg, vyg = rng.random(2 * P2), sim.vth * rng.standard_normal(2 * P2), sim.vth * rng.standard_normal(2 * P2)
    f, c = sim.first, sim.count
    sim.set_particles(xg[f:f + c], vxg[f:f + c], vy=vyg[f:f + c])
    sim.step(8)
    got = sim.particles(); fld = sim.fields(); D, _ = sim.diagnostics()
    ref = pg.ngp_1d2v_2s(N=256, P=P2, T=64, TO=16, M=8.0, device=local)
    ref.set_particles(xg, vxg, vy=vyg); ref.step(8)
    gr = ref.particles(); fr = ref.fields(); Dr, _ = ref.diagnostics()
    good = all(np.array_equal(a, b[f:f + c]) for a, b in zip(got, gr)) and np.array_equal(fld[0], fr[0]) and rel(D, Dr) < 1e-12
    print(f"[rank {rank}] ngp_1d2v_2s bit-identical={good}", flush=True)
    ok &= good
    sim.close(); ref.close()
    ok &= check_esfield(rank, world, local)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if int(t.item()) == 1 else "FAIL", flush=True)
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
