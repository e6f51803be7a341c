#!/usr/bin/env python3
"""Generate tests/golden/*.npz with the CPU oracle (oracle/picgolf_oracle.c).

The reference has no tests or golden vectors and Julia cannot run here, so these fixtures are
outputs of the oracle restatement on seeded inputs (inputs are stored too, so the fixtures do not
depend on numpy's RNG stream).  They pin (a) the oracle against accidental drift and (b) the CUDA
path on the GPU box, where neither /root/reference nor a long CPU run is available.

Run:  python tools/make_golden.py [--skip-c3]
"""
import argparse
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, os.path.getsize(path), "bytes")


def c1_ngp(steps=8):
    """Config 1: src/NGPFourier.jl  N=128, P=64N, dt=1/4N, W=200 (w=3.125 dyadic)."""
    N, P = 128, 64 * 128
    dt, W = 1 / (4 * N), 200.0
    w = W / P * N
    rng = np.random.default_rng(0)
    x0 = rng.random(P)
    v0 = np.where(np.arange(1, P + 1) > P / 2, 1.0, -1.0)
    x, v = x0.copy(), v0.copy()
    rhos, Es, raws = [], [], []
    for _ in range(steps):
        rho, E, raw = o.ngp_step(x, v, N, dt, w)
        rhos.append(rho); Es.append(E); raws.append(raw)
    save("c1_ngp", N=N, P=P, dt=dt, W=W, w=w, x0=x0, v0=v0, x=x, v=v, rho=np.array(rhos), E=np.array(Es),
         raw=np.array(raws), idx1=o.ngp_index(x0, N))


def gauss_explicit(steps=8):
    """src/Gaussian.jl  NX=128, NP=64NX, dt=1/10NX, W=1600, w=W/NP, deposit scale w/dx."""
    N, P = 128, 64 * 128
    dt, W = 1 / (10 * N), 1600.0
    scale = W / P / (1 / N)
    rng = np.random.default_rng(1)
    x0 = rng.random(P)
    v0 = np.where(np.arange(1, P + 1) > P / 2, 1.0, -1.0)
    x, v = x0.copy(), v0.copy()
    rhos, Es, raws = [], [], []
    for _ in range(steps):
        rho, E, raw = o.gauss_leapfrog_step(x, v, N, 6, dt, scale)
        rhos.append(rho); Es.append(E); raws.append(raw)
    save("gauss_explicit", N=N, P=P, dt=dt, W=W, scale=scale, x0=x0, v0=v0, x=x, v=v, rho=np.array(rhos),
         E=np.array(Es), raw=np.array(raws))


def c2_fixedpoint(steps=16):
    """Config 2: src/GaussianFixedPoint.jl  N=128, P=32N, dt=1/6N, W=400, +-6, l=1e-8."""
    N, P = 128, 32 * 128
    dt, W = 1 / (6 * N), 400.0
    rng = np.random.default_rng(2)
    x0 = rng.random(P)
    v0 = rng.choice([-1.0, 1.0], P)
    fp = o.FixedPoint(x0, v0, N, dt, W, hw=6, rtol=1e-8, atol=0.0)
    Ds, sws, Es, rhos = [], [], [], []
    x1 = v1 = None
    for t in range(steps):
        D4, raw, s = fp.step()
        Ds.append(D4); sws.append(s); Es.append(fp.E.copy()); rhos.append(fp.r.copy())
        if t == 0:
            x1, v1 = fp.x.copy(), fp.v.copy()
    save("c2_fixedpoint", N=N, P=P, dt=dt, W=W, w=fp.w, x0=x0, v0=v0, x1=x1, v1=v1, x=fp.x, v=fp.v, D=np.array(Ds),
         sweeps=np.array(sws, dtype=np.int32), E=np.array(Es), rho=np.array(rhos))


def c3_quiet(T=2 ** 13):
    """Config 3: src/GaussianFixedPointQuiet.jl  N=64, P=2048, T=2^13, W=32pi^2/3, +-7, l=4eps, quiet start."""
    N, P = 64, 2048
    dt, W = 1 / (6 * N), 32 * math.pi ** 2 / 3
    x0, v0 = o.quiet_start(P)
    fp = o.FixedPoint(x0, v0, N, dt, W, hw=7, rtol=4 * np.finfo(float).eps, atol=0.0)
    # first 16 steps one by one (state snapshots), then the rest in one call
    D, sw = np.zeros((T, 4)), np.zeros(T, dtype=np.int32)
    x16 = v16 = E16 = None
    for t in range(16):
        D[t], _, sw[t] = fp.step()
    x16, v16, E16 = fp.x.copy(), fp.v.copy(), fp.E.copy()
    Drest, swrest = fp.run(T - 16)
    D[16:], sw[16:] = Drest, swrest
    save("c3_quiet", N=N, P=P, dt=dt, W=W, w=fp.w, T=T, D=D, sweeps=sw, x16=x16, v16=v16, E16=E16,
         slope_pred=o.growth_slope(W))


def simpson13(T=2048):
    """src/GaussianFixedPointQuietSimpson13.jl (N=64, P=2048, +-7, l=4eps, quiet start), first
    2048 of its 2^13 steps, plus a noisy-start case (N=128, P=4096, +-6, l=1e-8) for step-level parity."""
    N, P = 64, 2048
    dt, W = 1 / (6 * N), 32 * math.pi ** 2 / 3
    x0, v0 = o.quiet_start(P)
    s = o.Simpson13(x0, v0, N, dt, W)
    D, sw = np.zeros((T, 4)), np.zeros(T, dtype=np.int32)
    for t in range(16):
        D[t], _, sw[t] = s.step()
    x16, v16, E16 = s.x.copy(), s.v.copy(), s.E.copy()
    Dr, swr = s.run(T - 16)
    D[16:], sw[16:] = Dr, swr
    rng = np.random.default_rng(13)
    xr, vr = rng.random(4096), rng.choice([-1.0, 1.0], 4096)
    n = o.Simpson13(xr, vr, 128, 1 / (6 * 128), 400.0, hw=6, rtol=1e-8)
    Dn, swn = [], []
    for t in range(8):
        d, _, k = n.step()
        Dn.append(d); swn.append(k)
    save("simpson13", N=N, P=P, dt=dt, W=W, T=T, D=D, sweeps=sw, x16=x16, v16=v16, E16=E16,
         xr=xr, vr=vr, xn=n.x, vn=n.v, En=n.E, rn=n.r, Dn=np.array(Dn), swn=np.array(swn, dtype=np.int32))


def area_simpson13(T=1024):
    """src/AreaFixedPointQuietSimpson13.jl (N=64, P=2048, area shape, l=1e-14, quiet start)."""
    N, P = 64, 2048
    dt, W = 1 / (6 * N), 32 * math.pi ** 2 / 3
    x0, v0 = o.quiet_start(P)
    s = o.Simpson13(x0, v0, N, dt, W, rtol=1e-14, shape=1)
    D, sw = np.zeros((T, 4)), np.zeros(T, dtype=np.int32)
    for t in range(16):
        D[t], _, sw[t] = s.step()
    x16, v16 = s.x.copy(), s.v.copy()
    Dr, swr = s.run(T - 16)
    D[16:], sw[16:] = Dr, swr
    rng = np.random.default_rng(17)
    xr, vr = rng.random(4096), rng.choice([-1.0, 1.0], 4096)
    n = o.Simpson13(xr, vr, 128, 1 / (6 * 128), 400.0, rtol=1e-9, shape=1)
    Dn, swn = [], []
    for t in range(8):
        d, _, k = n.step()
        Dn.append(d); swn.append(k)
    save("area_simpson13", N=N, P=P, dt=dt, W=W, T=T, D=D, sweeps=sw, x16=x16, v16=v16,
         xr=xr, vr=vr, xn=n.x, vn=n.v, En=n.E, rn=n.r, Dn=np.array(Dn), swn=np.array(swn, dtype=np.int32))


def ngp1d2v(steps=8):
    """src/NGP1D2V.jl (N=512, P=15N, erf shape +-7, Boris about z), inputs drawn like :27-32."""
    from scipy.special import erfinv
    N = 512
    P = 15 * N
    n0 = 4 * math.pi ** 2
    vth = math.sqrt(n0) / N / 4
    dt, B0, w = 1 / N / (6 * vth), math.sqrt(n0) / 16, n0 / P
    rng = np.random.default_rng(19)
    x0, vx0, vy0 = rng.random(P), vth * erfinv(rng.random(P)), vth * erfinv(rng.random(P))
    x, vx, vy = x0.copy(), vx0.copy(), vy0.copy()
    rhos, Es, raws = [], [], []
    for _ in range(steps):
        rho, E, raw = o.step_1d2v(x, vx, vy, N, 7, dt, B0, w)
        rhos.append(rho); Es.append(E); raws.append(raw)
    save("ngp1d2v", N=N, P=P, dt=dt, B0=B0, w=w, n0=n0, vth=vth, x0=x0, vx0=vx0, vy0=vy0, x=x, vx=vx, vy=vy,
         rho=np.array(rhos), E=np.array(Es), raw=np.array(raws))


def ngp1d2v2s(steps=8):
    """SURVEY 8f rank 2, two species: src/NGP1D2V2S.jl (N=256, P=8N per species, M=8); x as :17-18 (bit-reversed, second
    species shifted by pi), velocities drawn like NGP1D2V.jl:30-31 (the R(igr) sequences of :19-20 need Roots.jl)."""
    from scipy.special import erfinv
    N = 256
    P = 8 * N
    M = 8.0
    n0 = 4 * math.pi ** 2
    vth = math.sqrt(n0) / N / 8
    dt, B0, w = 1 / N / (16 * vth), math.sqrt(n0) / 8, n0 / (2 * P)
    rng = np.random.default_rng(23)
    xq, _ = o.quiet_start(P)
    x0 = np.concatenate([np.mod(xq, 1), np.mod(xq + math.pi, 1)])
    vx1, vy1 = vth * erfinv(rng.random(P)), vth * erfinv(rng.random(P))
    vx0, vy0 = np.concatenate([vx1, vx1 / math.sqrt(M)]), np.concatenate([vy1, vy1 / math.sqrt(M)])
    x, vx, vy = x0.copy(), vx0.copy(), vy0.copy()
    rhos, Es, raws = [], [], []
    for _ in range(steps):
        rho, E, raw = o.step_1d2v2s(x, vx, vy, N, 7, dt, B0, w, M)
        rhos.append(rho); Es.append(E); raws.append(raw)
    save("ngp1d2v2s", N=N, P=P, M=M, dt=dt, B0=B0, w=w, n0=n0, vth=vth, x0=x0, vx0=vx0, vy0=vy0, x=x, vx=vx, vy=vy,
         rho=np.array(rhos), E=np.array(Es), raw=np.array(raws))


def c5_2d3v(steps=4):
    """Config 5 shape at test size: src/Electrostatic2D3V.jl with NX=NY=32, P=NX*NY*8."""
    NX = NY = 32
    P = NX * NY * 8
    NG = math.sqrt(NX ** 2 + NY ** 2)
    n0 = 4 * math.pi ** 2
    vth = math.sqrt(n0) / NG
    dt = 1 / NG / (6 * vth)
    B0 = math.sqrt(n0) / 4
    w = n0 / P / ((1 / NX) * (1 / NY))
    rng = np.random.default_rng(5)
    x0, y0 = 1.0 - rng.random(P), 1.0 - rng.random(P)  # (0,1]
    vx0, vy0, vz0 = (rng.standard_normal(P) * vth / math.sqrt(2) for _ in range(3))
    x, y, vx, vy, vz = (a.copy() for a in (x0, y0, vx0, vy0, vz0))
    Ex, Ey = np.zeros(NX * NY), np.zeros(NX * NY)
    rhos, Exs, Eys, Ks = [], [], [], []
    for _ in range(steps):
        rho = o.step_2d3v(x, y, vx, vy, vz, NX, NY, dt, B0, w, Ex, Ey, nthreads=1)
        rhos.append(rho); Exs.append(Ex.copy()); Eys.append(Ey.copy())
        Ks.append(o.diagnostics_2d3v(Ex, Ey, NX, NY, vx, vy, w))
    save("c5_2d3v", NX=NX, NY=NY, P=P, dt=dt, B0=B0, w=w, n0=n0, vth=vth, x0=x0, y0=y0, vx0=vx0, vy0=vy0, vz0=vz0,
         x=x, y=y, vx=vx, vy=vy, vz=vz, rho=np.array(rhos), Ex=np.array(Exs), Ey=np.array(Eys), K=np.array(Ks))


def es_setup(NX=32, NY=16, ppc=6, Lx=2.0, Ly=1.0, shapes=(12, 13)):
    """Two-species magnetised plasma in the style of src/2D3V.jl:70-116 (electrons + ions of mass 16) at test size, on a
    non-square box with Lx != Ly, Halton starts of Species(...) (PIC2D3V.jl:194-213)."""
    P = NX * NY * ppc
    n0 = 4 * math.pi ** 2
    dl = min(Lx / NX, Ly / NY)
    vth = dl * math.sqrt(n0)
    dt = dl / (6 * vth)
    B = [math.sqrt(n0) / 4, math.sqrt(n0) / 16, -math.sqrt(n0) / 8]
    species = []
    for (q, m), sh in zip(((-1.0, 1.0), (1.0, 16.0)), shapes):
        x, y, vx, vy, vz, w = o.es_species(P, vth / math.sqrt(m), n0, Lx, Ly)
        species.append(dict(x=x, y=y, vx=vx, vy=vy, vz=vz, charge=q, mass=m, weight=w, shape=sh, vth=vth / math.sqrt(m)))
    return dict(NX=NX, NY=NY, P=P, Lx=Lx, Ly=Ly, n0=n0, dt=dt, B=B), species


def esfield(NT=12, ntskip=4, ngskip=2):
    """loop!/diagnose! of PIC2D3V.ElectrostaticField, both readings of update! (accumulate 1 = as written)."""
    par, species = es_setup()
    out = dict(NT=NT, ntskip=ntskip, ngskip=ngskip, **{k: np.asarray(v) for k, v in par.items()})
    for s, sp in enumerate(species):
        out[f"xyv0_{s}"] = np.stack([sp[k] for k in ("x", "y", "vx", "vy", "vz")], axis=1)  # P x 5 (Julia: xyv[5, P])
        out[f"spec_{s}"] = np.array([sp["charge"], sp["mass"], sp["weight"], sp["shape"], sp["vth"]])
    for acc in (1, 0):
        f = o.ESField(species, par["NX"], par["NY"], par["Lx"], par["Ly"], par["dt"], par["B"], NT=NT, ntskip=ntskip, ngskip=ngskip,
                      accumulate=bool(acc))
        rhos, Exs_, Eys_ = [], [], []
        for _ in range(NT):
            f.step()
            rhos.append(f.rho.copy()); Exs_.append(f.Ex.copy()); Eys_.append(f.Ey.copy())
        gx, gy = f.exy_interior()
        tag = f"acc{acc}_"
        out.update({tag + "rho": np.array(rhos), tag + "Ex": np.array(Exs_), tag + "Ey": np.array(Eys_), tag + "Exy_x": gx, tag + "Exy_y": gy,
                    tag + "scalars": f.scalars.copy(), tag + "Exs": f.Exs.copy(), tag + "Eys": f.Eys.copy(), tag + "phis": f.phis.copy(),
                    tag + "x": f.x.copy(), tag + "y": f.y.copy(), tag + "vx": f.vx.copy(), tag + "vy": f.vy.copy(), tag + "vz": f.vz.copy()})
    save("esfield", **out)


def stencils():
    """Stage-level vectors: Gaussian stencils at awkward centres for N = 64, 128, 4096."""
    rng = np.random.default_rng(7)
    out = {}
    for N, hw in ((64, 7), (128, 6), (4096, 6), (4096, 7)):
        c = np.concatenate([rng.random(256), rng.random(32) * 0.02 - 0.01, 1 + rng.random(32) * 0.02 - 0.01,
                            (np.arange(8) + 0.5) / N, [0.0, 1.0, 0.5, 1e-300, -1e-300]])
        idx = np.zeros((c.size, 2 * hw + 1), dtype=np.int32)
        wt = np.zeros((c.size, 2 * hw + 1))
        for j, cj in enumerate(c):
            idx[j], wt[j] = o.gauss_stencil(cj, N, hw)
        out[f"c_{N}_{hw}"], out[f"idx_{N}_{hw}"], out[f"wt_{N}_{hw}"] = c, idx, wt
    save("stencils", **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-c3", action="store_true")
    ap.add_argument("--only", default=None, help="generate a single fixture, e.g. simpson13")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    if args.only:
        globals()[args.only]()
        sys.exit(0)
    c1_ngp(); gauss_explicit(); c2_fixedpoint(); c5_2d3v(); stencils(); ngp1d2v(); ngp1d2v2s(); esfield()
    if not args.skip_c3:
        c3_quiet()
        simpson13()
        area_simpson13()
