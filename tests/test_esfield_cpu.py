"""CPU checks of the PIC2D3V.jl electrostatic path: known answers that pin the oracle restatement
(oracle_es_*), the committed golden fixture, and -- through tests/es_host_harness.cpp -- the arithmetic of the header the
CUDA kernels evaluate on the device (pg_es_math.h) plus the periodic-grid / fixed-point design of es_particles_kernel
against the oracle's literal halo-array version.  No GPU, no compute call into libpicgolf.so."""
import ctypes as C
import math
import os
import subprocess
import sys

import numpy as np
import pytest
from conftest import ROOT, golden, relnorm

CSRC = os.path.join(ROOT, "particleincellcodegolf.jl_b200", "csrc")
SHAPES = [0, 1, 10, 11, 12, 13, 14, 15]
SUPPORT = {0: 1, 1: 2, 10: 1, 11: 2, 12: 3, 13: 4, 14: 5, 15: 6}


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("es_harness") / "es_host_harness.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", CSRC,
                           os.path.join(ROOT, "tests", "es_host_harness.cpp"), "-o", so])
    L = C.CDLL(so)
    dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.h_shape.restype = C.c_int
    L.h_shape.argtypes = [C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int), dp]
    L.h_boris.argtypes = [dp, C.c_double, C.c_double, dp, C.c_double, C.c_double]
    L.h_halton.restype = C.c_double
    L.h_halton.argtypes = [C.c_longlong, C.c_int, C.c_double]
    L.h_emulate_particles.argtypes = [C.c_int, C.c_longlong, dp, dp, dp, dp, dp, dp, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"),
                                      C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, dp, C.c_double, C.c_double, C.c_double, dp]
    return L


# ---------------------------------------------------------------- oracle known answers
def test_unimod_and_halton(oracle):
    # halton(i, base, seed): PIC2D3V.jl:29-37; sample(P, i) = halton.(0:P-1, i, 1/sqrt(2))  :191
    seed = 1 / math.sqrt(2)
    assert oracle.es_halton(0, 2, seed) == seed
    assert oracle.es_halton(1, 2, 0.0) == 0.5 and oracle.es_halton(2, 2, 0.0) == 0.25 and oracle.es_halton(3, 2, 0.0) == 0.75
    assert oracle.es_halton(1, 3, 0.0) == 1 / 3 and oracle.es_halton(5, 3, 0.0) == pytest.approx(2 / 3 + 1 / 9, abs=1e-16)
    assert oracle.es_halton(1, 2, seed) == (0.5 + seed) - 1.0


@pytest.mark.parametrize("shape", SHAPES)
def test_shape_fractions(oracle, shape):
    """Fractions sum to 1, match the cardinal B-spline of that order at the grid points, and sit on the right cells."""
    from scipy.interpolate import BSpline
    rng = np.random.default_rng(shape)
    NZ_Lz = 64 / 1.5
    for z in np.concatenate([rng.random(300) * 1.5, [1e-9, 1.5 - 1e-12, 0.75]]):
        j0, w = oracle.es_shape(shape, z, NZ_Lz)
        assert len(w) == SUPPORT[shape]
        assert abs(w.sum() - 1) < 2e-14
        zc = z * NZ_Lz  # position in cells
        r = math.ceil(zc) - zc
        if shape == 0:
            assert j0 == math.ceil(zc)
            continue
        order = 1 if shape == 1 else shape - 10
        assert j0 == math.ceil(zc) + (1 if (order % 2 == 0 and r > 0.5) else 0) - order // 2
        # Independent ground truth for the polynomial coefficients: the fractions are the values of the cardinal B-spline
        # of that degree on a unit lattice.  The reference's index convention (weight 1-r on cell i, r on cell i+1 for the
        # area shape, src/PIC2D3V.jl:1117-1119, and its mirror images for the higher orders) fixes the lattice phase up to
        # a reflection, so compare as sets over the admissible phases.
        b = BSpline.basis_element(np.arange(order + 2) - (order + 1) / 2)
        best = np.inf
        for phase in (r, 1 - r, r + 0.5, 0.5 - r, r - 0.5, 1.5 - r):
            vals = np.nan_to_num(b(phase % 1.0 + np.arange(-order - 1, order + 2), extrapolate=False))
            vals = np.sort(vals[vals > 0])[::-1][: order + 1]
            ws = np.sort(w)[::-1]
            if len(vals) == len(ws):
                best = min(best, np.abs(vals - ws).max())
            elif len(vals) < len(ws):  # a weight that is exactly zero at a lattice point
                best = min(best, np.abs(np.concatenate([vals, np.zeros(len(ws) - len(vals))]) - ws).max())
        assert best < 5e-14, (shape, z, w)


def test_area_is_bspline1_and_ngp_index(oracle):
    rng = np.random.default_rng(1)
    for z in rng.random(200):
        ja, wa = oracle.es_shape(1, z, 32.0)
        jb, wb = oracle.es_shape(11, z, 32.0)
        assert ja == jb and np.array_equal(wa, wb)  # (i, 1-r), (i+1, r) == bspline{1}(1 - r) on i:(i+1)
        j0, w0 = oracle.es_shape(10, z, 32.0)      # BSpline{0}: the nearer of cells i, i+1 by r > 0.5
        r = math.ceil(z * 32.0) - z * 32.0
        assert j0 == math.ceil(z * 32.0) + (1 if r > 0.5 else 0) and w0[0] == 1.0


def test_boris_known_answers(oracle):
    v0 = np.array([0.3, -0.2, 0.5])
    B, dt = [0.4, -0.7, 0.2], 0.05
    for q_m in (1.0, -1.0):  # |q_m| = 1: a pure rotation when E = 0
        v = oracle.es_boris(v0, 0.0, 0.0, B, dt, q_m)
        assert abs(np.linalg.norm(v) - np.linalg.norm(v0)) < 1e-15
    # B = 0: v += E*dt*q_m exactly as two half kicks
    v = oracle.es_boris(v0, 2.0, -3.0, [0, 0, 0], dt, 0.5)
    assert np.allclose(v, v0 + np.array([2.0, -3.0, 0.0]) * dt * 0.5, rtol=0, atol=1e-16)
    # the push as written rotates about the UNSCALED t = B dt/2 with the factor q_m^2 * 2 / (1 + q_m^2 t^2)   :49-54
    q_m = 1 / 16
    t = np.array(B) * dt / 2
    vm = v0.copy()
    want = vm + np.cross(vm + np.cross(vm, t), t) * q_m ** 2 * 2 / (1 + q_m ** 2 * t.dot(t))
    assert np.allclose(oracle.es_boris(v0, 0.0, 0.0, B, dt, q_m), want, rtol=0, atol=1e-16)


def _one_species(oracle, shape, NX=16, NY=16, ppc=4, Lx=1.0, Ly=1.0, charge=-1.0, mass=1.0, seed=0):
    rng = np.random.default_rng(seed)
    P = NX * NY * ppc
    vth = 0.02
    return dict(x=Lx * (1 - rng.random(P)), y=Ly * (1 - rng.random(P)), vx=rng.standard_normal(P) * vth, vy=rng.standard_normal(P) * vth,
                vz=rng.standard_normal(P) * vth, charge=charge, mass=mass, weight=4 * math.pi ** 2 * Lx * Ly / P, shape=shape)


def test_single_mode_solve_with_box_lengths(oracle):
    """rho = A cos(2 pi m x / Lx) placed by NGP particles sitting on cell centres -> Ex = A sin(.) / k, Ey = 0."""
    NX, NY, Lx, Ly, m = 32, 16, 2.5, 0.7, 3
    ii, jj = np.meshgrid(np.arange(1, NX + 1), np.arange(1, NY + 1), indexing="ij")
    xc, yc = (ii - 0.5) * Lx / NX, (jj - 0.5) * Ly / NY
    k = 2 * math.pi * m / Lx
    cnt = (8 + np.rint(4 * np.cos(k * xc))).astype(int)  # that many NGP particles on every cell centre
    x = np.repeat(xc.ravel(order="F"), cnt.ravel(order="F"))
    y = np.repeat(yc.ravel(order="F"), cnt.ravel(order="F"))
    sp = dict(x=x, y=y, vx=np.zeros(x.size), vy=np.zeros(x.size), vz=np.zeros(x.size), charge=1.0, mass=1.0, weight=1.0, shape=0)
    f = oracle.ESField([sp], NX, NY, Lx, Ly, 1e-3, [0, 0, 0], NT=1)
    f.step()
    dV = (Lx / NX) * (Ly / NY)
    rho = f.rho.reshape(NY, NX).T
    assert np.allclose(rho, cnt / dV, rtol=1e-14, atol=0)
    # differentiate the deposited density spectrally with numpy: Ex^ = -i kx rho^ / k^2, DC dropped; real() drops Nyquist
    rk = np.fft.fft2(rho)
    kx = 2 * math.pi / Lx * np.fft.fftfreq(NX, 1 / NX)[:, None]
    ky = 2 * math.pi / Ly * np.fft.fftfreq(NY, 1 / NY)[None, :]
    k2 = kx ** 2 + ky ** 2
    k2[0, 0] = 1.0
    exk = -1j * kx * rk / k2
    exk[0, 0] = 0.0
    want = np.real(np.fft.ifft2(exk))
    got = f.Ex.reshape(NY, NX).T
    assert relnorm(got, want) < 1e-12
    assert np.abs(f.Ey).max() < 1e-12 * np.abs(got).max()
    # and the dominant mode is the analytic one: rho ~ (4/dV) cos(k x) -> Ex ~ (4/dV) sin(k x) / k (rounded counts: few %)
    ana = 4 / dV * np.sin(k * xc) / k
    assert relnorm(got, ana) < 0.1


def test_thread_chunks_only_change_rounding(oracle):
    sp = [_one_species(oracle, 12, seed=1), _one_species(oracle, 13, charge=1.0, mass=4.0, seed=2)]
    runs = []
    for nth in (1, 3):
        f = oracle.ESField(sp, 16, 16, 1.0, 1.0, 0.01, [1.0, 0.3, -0.2], NT=4, nthreads=nth)
        for _ in range(4):
            f.step()
        runs.append(f)
    assert relnorm(runs[1].rho, runs[0].rho) < 1e-12 and relnorm(runs[1].x, runs[0].x) < 1e-13


def test_update_accumulates_as_written(oracle):
    """update! adds the new field to Exy and nothing zeroes it (PIC2D3V.jl:294-297, 21-27): Exy = running sum of E."""
    sp = [_one_species(oracle, 1, seed=3)]
    f = oracle.ESField(sp, 16, 16, 1.0, 1.0, 0.01, [0.5, 0, 0], NT=3, accumulate=True)
    tot = np.zeros(256)
    for _ in range(3):
        f.step()
        tot += f.Ex
    assert np.array_equal(f.exy_interior()[0], tot)
    g = oracle.ESField(sp, 16, 16, 1.0, 1.0, 0.01, [0.5, 0, 0], NT=3, accumulate=False)
    for _ in range(3):
        g.step()
    assert np.array_equal(g.exy_interior()[0], g.Ex)


def test_field_energy_counts_the_halo(oracle):
    """mean(abs2, f.Exy)/2 runs over the halo array: cells within 3 of an edge count twice per dimension."""
    NX, NY = 16, 32
    f = oracle.ESField([_one_species(oracle, 1, NX=NX, NY=NY, seed=4)], NX, NY, 1.0, 1.0, 0.01, [0, 0, 0], NT=2)
    f.step(); f.step()
    gx, gy = f.exy_interior()
    mult = lambda n: np.array([1 + (i <= 3) + (i > n - 3) for i in range(1, n + 1)])
    m2 = np.outer(mult(NX), mult(NY)).ravel(order="F")
    want = (m2 * (gx ** 2 + gy ** 2)).sum() / (2 * (NX + 6) * (NY + 6)) / 2
    assert f.scalars[1, 1] == pytest.approx(want, rel=1e-13)


def test_momentum_is_conserved_without_B(oracle):
    """Same shape for gather and deposit, spectral solve: total particle momentum is conserved to round-off."""
    sp = [_one_species(oracle, 12, seed=5), _one_species(oracle, 12, charge=1.0, mass=9.0, seed=6)]
    f = oracle.ESField(sp, 16, 16, 1.0, 1.0, 0.02, [0, 0, 0], NT=8, accumulate=False)
    for _ in range(8):
        f.step()
    p = f.scalars[:, 2:5]
    c = f.scalars[:, 5:8]
    assert np.abs(p - p[0]).max() < 1e-13 * c.max()


def test_golden_fixture_is_reproduced(oracle):
    g = golden("esfield")
    species = []
    for s in range(2):
        a, spec = g[f"xyv0_{s}"], g[f"spec_{s}"]
        species.append(dict(x=a[:, 0], y=a[:, 1], vx=a[:, 2], vy=a[:, 3], vz=a[:, 4], charge=spec[0], mass=spec[1], weight=spec[2],
                            shape=int(spec[3])))
    for acc in (1, 0):
        f = oracle.ESField(species, int(g["NX"]), int(g["NY"]), float(g["Lx"]), float(g["Ly"]), float(g["dt"]), g["B"], NT=int(g["NT"]),
                           ntskip=int(g["ntskip"]), ngskip=int(g["ngskip"]), accumulate=bool(acc))
        for _ in range(int(g["NT"])):
            f.step()
        t = f"acc{acc}_"
        assert np.array_equal(f.x, g[t + "x"]) and np.array_equal(f.vz, g[t + "vz"])
        assert np.array_equal(f.scalars, g[t + "scalars"]) and np.array_equal(f.Exs, g[t + "Exs"]) and np.array_equal(f.phis, g[t + "phis"])
        # "phis" is rho - mean(rho) (phi holds the spectrum of rho with [1,1] zeroed): what the kernels store directly
        rho = g[t + "rho"]
        ntskip, gs = int(g["ntskip"]), int(g["ngskip"])
        NX, NY = int(g["NX"]), int(g["NY"])
        want = np.zeros_like(f.phis)
        for step in range(int(g["NT"])):
            r = rho[step].reshape(NY, NX).T
            want[:, :, step // ntskip] += (r - r.mean())[::gs, ::gs] / ntskip
        assert relnorm(f.phis, want) < 1e-13
    # the Halton start of the fixture is the oracle's Species(...)
    x, y, vx, vy, vz, w = oracle.es_species(int(g["P"]), float(g["spec_0"][4]), float(g["n0"]), float(g["Lx"]), float(g["Ly"]))
    assert np.array_equal(x, g["xyv0_0"][:, 0]) and np.array_equal(vz, g["xyv0_0"][:, 4]) and w == g["spec_0"][2]
    assert abs(vx.mean()) < 1e-18 and vx.std(ddof=1) == pytest.approx(float(g["spec_0"][4]) / math.sqrt(2), rel=1e-14)


# ---------------------------------------------------------------- the device header on the host
@pytest.mark.parametrize("shape", SHAPES)
def test_device_header_shapes_bit_exact(oracle, harness, shape):
    rng = np.random.default_rng(100 + shape)
    w6 = np.zeros(6)
    for z in np.concatenate([rng.random(500) * 3.0, [1e-12, 3.0, 1.5]]):
        j = C.c_int()
        n = harness.h_shape(shape, float(z), 128 / 3.0, C.byref(j), w6)
        j0, w = oracle.es_shape(shape, z, 128 / 3.0)
        assert n == len(w) and j.value == j0 and np.array_equal(w6[:n], w)


def test_device_header_boris_and_halton_bit_exact(oracle, harness):
    rng = np.random.default_rng(7)
    # (the header divides by the species-wide 1 + q_m^2 t2 through its precomputed reciprocal and an FMA correction: must be the
    # same bits as the oracle's IEEE division for any charge-to-mass ratio, field strength and velocity scale)
    for n in range(20000):
        v = rng.standard_normal(3) * 10.0 ** rng.integers(-6, 7)
        B = rng.standard_normal(3) * 10.0 ** rng.integers(-3, 4)
        q_m = rng.choice([-1.0, 1.0, 1 / 16, -0.25]) if n % 2 else rng.standard_normal() * 10.0 ** rng.integers(-2, 3)
        ex, ey, dt = rng.standard_normal(), rng.standard_normal(), 0.01 + rng.random() * 0.1
        a = v.copy()
        harness.h_boris(a, ex, ey, np.ascontiguousarray(B), dt, q_m)
        assert np.array_equal(a, oracle.es_boris(v, ex, ey, B, dt, q_m))
    seed = 1 / math.sqrt(2)
    for base in (2, 3, 5, 7, 9):
        for i in list(range(64)) + [10 ** 6 + 7, 2 ** 40 + 1]:
            assert harness.h_halton(i, base, seed) == oracle.es_halton(i, base, seed)


@pytest.mark.parametrize("shapes", [(0, 1), (12, 13), (14, 15), (10, 11)])
def test_periodic_grid_emulation_matches_halo_oracle(oracle, harness, shapes):
    """es_particles_kernel's design (one periodic grid, unimod-wrapped indices, fixed-point charge in units of wref) run
    serially on the host reproduces the oracle's halo-array loop! for rho, the particles and the diagnostics sums."""
    NX, NY, Lx, Ly, dt = 16, 32, 1.5, 2.0, 0.02
    B = np.array([0.8, -0.3, 0.5])
    sp = [_one_species(oracle, shapes[0], NX=NX, NY=NY, Lx=Lx, Ly=Ly, seed=8), _one_species(oracle, shapes[1], NX=NX, NY=NY, Lx=Lx, Ly=Ly,
                                                                                            charge=2.0, mass=5.0, seed=9)]
    sp[1]["weight"] *= 0.5
    f = oracle.ESField(sp, NX, NY, Lx, Ly, dt, B, NT=8, accumulate=True)
    for _ in range(3):
        f.step()
    # state before step 4
    gx, gy = f.exy_interior()
    Exy2 = np.ascontiguousarray(np.stack([gx, gy], axis=1).ravel())
    parts = [a.copy() for a in (f.x, f.y, f.vx, f.vy, f.vz)]
    f.step()
    dV = (Lx / NX) * (Ly / NY)
    qw = [s["charge"] * s["weight"] / dV for s in sp]
    wref = max(abs(q) for q in qw)
    total = sum(len(s["x"]) for s in sp)
    frac = max(8, min(60, 62 - math.ceil(math.log2(total + 1))))
    rho_fx = np.zeros(NX * NY, dtype=np.int64)
    base, sums = 0, []
    for s, spec in enumerate(sp):
        P = len(spec["x"])
        loc = [np.ascontiguousarray(a[base:base + P]) for a in parts]
        sm = np.zeros(7)
        harness.h_emulate_particles(spec["shape"], P, *loc, Exy2, rho_fx, NX, NY, Lx, Ly, dt, B, spec["charge"] / spec["mass"],
                                    qw[s] / wref, float(2 ** frac), sm)
        for a, b in zip(loc, (f.x, f.y, f.vx, f.vy, f.vz)):
            assert np.array_equal(a, b[base:base + P])  # same expressions, same bits
        sums.append(sm)
        base += P
    rho = rho_fx.astype(np.float64) * 2.0 ** -frac * wref
    assert relnorm(rho, f.rho) < 1e-12
    ke = sum(sm[0] * spec["mass"] / 2 * spec["weight"] for sm, spec in zip(sums, sp))
    out = np.zeros(8)
    oracle.lib().oracle_es_diagnose(2, f.sP, f.smass, f.sweight, f.vx, f.vy, f.vz, NX, NY, f.Exy, out)
    assert ke == pytest.approx(out[0], rel=1e-13)
