"""N>1 host logic on CPU: world_size-2 gloo.  Particles shard by contiguous global index range, each
rank deposits into a full local grid, rho is summed with an all-reduce and the field solve runs
redundantly (CPU analogue src/Electrostatic2D3V.jl:114,126-141).  The per-rank compute
is stood in for by the oracle here (the CUDA path needs a GPU); what is under test is the sharding
rule of picgolf_create / shard_range, the quiet start by global index, the id broadcast helper and
that sharded == unsharded within summation-order tolerance."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import relnorm


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import particleincellcodegolf.jl_b200 as pg
        from particleincellcodegolf.jl_b200 import distributed as pgd
        from oracle import oracle as o

        # 1. id broadcast helper (what connect() uses to ship the NCCL unique id)
        payload = bytes(range(128)) if rank == 0 else None
        got = pgd.broadcast_bytes(payload, 128, src=0)
        assert got == bytes(range(128))

        # 2. quiet start by global index: shard == slice of the global population
        N, P, hw = 64, 2048 + 3, 7  # odd P: remainder goes to the low ranks
        first, count = pg.shard_range(P, rank, world)
        xg, vg = o.quiet_start(P)
        xs, vs = o.quiet_start(P, first, count)
        assert np.array_equal(xs, xg[first:first + count]) and np.array_equal(vs, vg[first:first + count])

        # 3. Gaussian deposit: sum of shard grids == unsharded grid up to summation order
        W = 32 * np.pi ** 2 / 3
        w = W / P * N
        r_loc = torch.from_numpy(o.gauss_deposit(xs, xs, N, hw, w))
        dist.all_reduce(r_loc)
        r_all = o.gauss_deposit(xg, xg, N, hw, w)
        assert relnorm(r_loc.numpy(), r_all) < 1e-13
        # all ranks hold identical bits after the all-reduce -> identical E, identical sweep decisions
        gathered = [torch.zeros_like(r_loc) for _ in range(world)]
        dist.all_gather(gathered, r_loc)
        assert all(torch.equal(gathered[0], g) for g in gathered)

        # 4. NGP deposit with dyadic w is bit-identical however it is sharded
        rng = np.random.default_rng(11)
        x = rng.random(8192)
        f, c = pg.shard_range(8192, rank, world)
        n_loc = torch.from_numpy(o.ngp_deposit(x[f:f + c], 128, 3.125))
        dist.all_reduce(n_loc)
        assert np.array_equal(n_loc.numpy(), o.ngp_deposit(x, 128, 3.125))

        # 5. one sharded fixed-point step == unsharded step (sweep counts equal; x, v within 1e-12)
        N2, P2 = 128, 4096
        x0 = rng.random(P2)
        v0 = rng.choice([-1.0, 1.0], P2)
        dt, W2 = 1 / (6 * N2), 400.0
        w2 = W2 / P2 * N2
        ref = o.FixedPoint(x0, v0, N2, dt, W2, hw=6, rtol=1e-8)
        _, _, s_ref = ref.step()
        f, c = pg.shard_range(P2, rank, world)
        X, V = x0[f:f + c].copy(), v0[f:f + c].copy()
        xx, vv = X.copy(), V.copy()
        E, F = np.zeros(N2), np.full(N2, np.nan)
        sweeps = 0
        for _ in range(10):
            if o.isapprox(F, E, 1e-8):
                break
            F = E.copy()
            xx = X + (vv + V) / 2 * dt
            r = torch.from_numpy(o.gauss_deposit(xx, X, N2, 6, w2))
            dist.all_reduce(r)
            E = o.solve1d(r.numpy())
            vv = V + o.gauss_gather(E, (xx + X) / 2, N2, 6) * dt
            sweeps += 1
        assert sweeps == s_ref == 4
        assert relnorm(E, ref.E) < 1e-12
        assert relnorm(vv, ref.v[f:f + c]) < 1e-12

        # 6. PIC2D3V species (include/picgolf_es.h): the Halton start comes from the GLOBAL particle index, its mean and
        #    corrected variance from all-reduced shard sums (what picgolf_es_init_species does with NCCL)
        from scipy.special import erfinv
        Ps, vth, seed = 1024 + 5, 0.013, 1 / np.sqrt(2.0)
        f, c = pg.shard_range(Ps, rank, world)
        xo, yo, vxo, vyo, vzo, wgt = o.es_species(Ps, vth, 4 * np.pi ** 2, 2.0, 0.5)
        idx = np.arange(f, f + c)
        assert np.array_equal(2.0 * np.array([o.es_halton(i, 2, seed) for i in idx]), xo[f:f + c])
        raw = vth * erfinv(2 * np.array([o.es_halton(i, 5, seed) for i in idx]) - 1) * vth

        def gsum(a):
            t = torch.tensor([float(np.sum(a))], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t.item())

        v = raw - gsum(raw) / Ps
        m2 = gsum(v) / Ps
        sd = np.sqrt(gsum((v - m2) ** 2) / (Ps - 1))
        v = v * ((vth / np.sqrt(2.0)) / sd)
        assert relnorm(v, vxo[f:f + c]) < 1e-12

        # 7. one sharded step of loop! (first step: Exy = 0): the shard grids sum to the unsharded charge density
        NXe, NYe, Lxe, Lye = 16, 32, 2.0, 0.5
        full = dict(x=xo, y=yo, vx=vxo, vy=vyo, vz=vzo, charge=-1.0, mass=1.0, weight=wgt, shape=12)
        mine = {k: (a[f:f + c].copy() if isinstance(a, np.ndarray) else a) for k, a in full.items()}
        fa = o.ESField([full], NXe, NYe, Lxe, Lye, 0.01, [0.3, 0.0, 0.1], NT=1)
        fb = o.ESField([mine], NXe, NYe, Lxe, Lye, 0.01, [0.3, 0.0, 0.1], NT=1)
        fa.step(); fb.step()
        r = torch.from_numpy(fb.rho.copy())
        dist.all_reduce(r)
        assert relnorm(r.numpy(), fa.rho) < 1e-13
        assert np.array_equal(fb.x, fa.x[f:f + c]) and np.array_equal(fb.vz, fa.vz[f:f + c])
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        out.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers(pg):
    for P in (1, 7, 2048, 2 ** 28 + 5):
        for n in (1, 2, 3, 8):
            pos = 0
            for r in range(n):
                f, c = pg.shard_range(P, r, n)
                assert f == pos and c >= 0
                pos += c
            assert pos == P


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(m == "ok" for _, m in res), res
