"""GPU parity of the PIC2D3V.jl electrostatic path and the omega-k post-processing (rank 4), all
through the C ABI of include/picgolf_es.h, against the oracle restatement (oracle_es_*) and the golden fixture.
Bars: shape fractions, Boris, xyv round trip, repeated runs: bit-exact; rho, E, particles after several steps:
|d| <= 1e-11 max|.| (fixed-point charge in a different summation order than the reference's thread grids); scalars 1e-11;
spectra 1e-10 against numpy's FFT."""
import math
import os

import numpy as np
import pytest
from conftest import golden, relnorm

pytestmark = pytest.mark.gpu

SHAPES = [0, 1, 10, 11, 12, 13, 14, 15]
TOL = 1e-12
TOL_TILED = 3e-11  # tile-sorted vs any-order kernel: two fixed-point formats (per deposit / per window), measured against the NET charge of two
                   # opposite species (a few % of one species' density): measured 2e-12 ... 9e-12, see PARITY.json


@pytest.fixture(scope="module")
def es(pg):
    import particleincellcodegolf.jl_b200.pic2d3v as m
    return m


def _shape(es, code):
    return es.NGPWeighting() if code == 0 else es.AreaWeighting() if code == 1 else es.BSplineWeighting(code - 10)


def _species_from_arrays(es, a, charge, mass, weight, shape, Lx, Ly):
    sp = es.Species(a.shape[0], 1.0, 1.0, _shape(es, shape), Lx=Lx, Ly=Ly, charge=charge, mass=mass, xyv=a)
    sp.weight = float(weight)
    return sp


@pytest.mark.parametrize("shape", SHAPES)
def test_stage_shape_bit_exact(es, oracle, shape):
    rng = np.random.default_rng(shape)
    z = np.concatenate([rng.random(2000) * 3.0, [1e-12, 3.0, 1.5, 0.75]])
    j0, wt = es.shape_weights(shape, z, 128 / 3.0)
    for k, zk in enumerate(z):
        jo, wo = oracle.es_shape(shape, zk, 128 / 3.0)
        assert j0[k] == jo and np.array_equal(wt[k, :len(wo)], wo) and not wt[k, len(wo):].any()


def test_stage_boris_bit_exact(es, oracle):
    rng = np.random.default_rng(3)
    n = 500
    v = rng.standard_normal((3, n))
    ex, ey = rng.standard_normal(n), rng.standard_normal(n)
    B, dt = [0.4, -0.7, 0.2], 0.03
    for q_m in (1.0, -1.0, 1 / 16):
        a, b, c = es.boris(v[0], v[1], v[2], ex, ey, B, dt, q_m)
        for k in range(n):
            assert np.array_equal([a[k], b[k], c[k]], oracle.es_boris(v[:, k], ex[k], ey[k], B, dt, q_m))


def _run_against(es, oracle, species_o, NX, NY, Lx, Ly, dt, B, NT, ntskip, ngskip, acc, sort_every=-1):
    f = oracle.ESField(species_o, NX, NY, Lx, Ly, dt, B, NT=NT, ntskip=ntskip, ngskip=ngskip, accumulate=bool(acc))
    plasma = [_species_from_arrays(es, np.stack([s[k] for k in ("x", "y", "vx", "vy", "vz")], axis=1), s["charge"], s["mass"], s["weight"],
                                   s["shape"], Lx, Ly) for s in species_o]
    sim = es.Simulation(plasma, es.ElectrostaticField(NX, NY, Lx, Ly, dt=dt, B0x=B[0], B0y=B[1], B0z=B[2], accumulate=bool(acc)),
                        es.ElectrostaticDiagnostics(NX, NY, NT, ntskip, ngskip), sort_every=sort_every)
    for t in range(NT):
        f.step()
        sim.loop(1)
        if t in (0, NT - 1):
            fl = sim.fields()
            assert relnorm(fl["rho"].ravel(order="F"), f.rho) < TOL, ("rho", t)
            assert relnorm(fl["Ex"].ravel(order="F"), f.Ex) < TOL and relnorm(fl["Ey"].ravel(order="F"), f.Ey) < TOL, ("E", t)
    gx, gy = f.exy_interior()
    fl = sim.fields()
    assert relnorm(fl["Exy_x"].ravel(order="F"), gx) < TOL and relnorm(fl["Exy_y"].ravel(order="F"), gy) < TOL
    base = 0
    for s, sp in enumerate(species_o):
        P = len(sp["x"])
        got = sim.species(s)
        for a, b, scale in zip(got, (f.x, f.y, f.vx, f.vy, f.vz), (Lx, Ly, None, None, None)):
            ref = b[base:base + P]
            d = np.abs(a - ref)
            if scale is not None:
                d = np.minimum(d, scale - d)  # a particle may sit on either side of the periodic seam
            assert d.max() <= TOL * max(np.abs(ref).max(), 1e-300)
        base += P
    sc = sim.scalars()
    rows = len(sc["kineticenergy"])
    assert rows == min(NT // ntskip, (NT - 1) // ntskip + 1)
    assert relnorm(sc["kineticenergy"], f.scalars[:rows, 0]) < TOL and relnorm(sc["fieldenergy"], f.scalars[:rows, 1]) < TOL
    cm = f.scalars[:rows, 5:8]
    assert np.abs(sc["characteristicmomentum"] - cm).max() < TOL * cm.max()
    assert np.abs(sc["particlemomentum"] - f.scalars[:rows, 2:5]).max() < TOL * cm.max()
    for name, ref in (("Exs", f.Exs), ("Eys", f.Eys), ("phis", f.phis)):
        h = sim.history(name)
        assert h.shape == ref[:, :, :rows].shape and relnorm(h, ref[:, :, :rows]) < TOL, name
    return sim, f


@pytest.mark.parametrize("acc", [1, 0])
def test_golden_fixture(es, oracle, acc):
    g = golden("esfield")
    NX, NY, Lx, Ly, dt, NT = int(g["NX"]), int(g["NY"]), float(g["Lx"]), float(g["Ly"]), float(g["dt"]), int(g["NT"])
    plasma = []
    for s in range(2):
        a, spec = g[f"xyv0_{s}"], g[f"spec_{s}"]
        plasma.append(_species_from_arrays(es, a, spec[0], spec[1], spec[2], int(spec[3]), Lx, Ly))
    B = g["B"]
    sim = es.Simulation(plasma, es.ElectrostaticField(NX, NY, Lx, Ly, dt=dt, B0x=B[0], B0y=B[1], B0z=B[2], accumulate=bool(acc)),
                        es.ElectrostaticDiagnostics(NX, NY, NT, int(g["ntskip"]), int(g["ngskip"])))
    sim.loop(NT)
    t = f"acc{acc}_"
    fl = sim.fields()
    assert relnorm(fl["rho"].ravel(order="F"), g[t + "rho"][-1]) < TOL
    assert relnorm(fl["Ex"].ravel(order="F"), g[t + "Ex"][-1]) < TOL and relnorm(fl["Ey"].ravel(order="F"), g[t + "Ey"][-1]) < TOL
    assert relnorm(fl["Exy_x"].ravel(order="F"), g[t + "Exy_x"]) < TOL
    P = int(g["P"])
    for s in range(2):
        x, y, vx, vy, vz = sim.species(s)
        assert relnorm(vx, g[t + "vx"][s * P:(s + 1) * P]) < TOL and relnorm(vz, g[t + "vz"][s * P:(s + 1) * P]) < TOL
        d = np.abs(x - g[t + "x"][s * P:(s + 1) * P])
        assert np.minimum(d, Lx - d).max() < TOL * Lx
    sc = sim.scalars()
    ref = g[t + "scalars"]
    assert relnorm(sc["kineticenergy"], ref[:, 0]) < TOL and relnorm(sc["fieldenergy"], ref[:, 1]) < TOL
    assert relnorm(sim.history("Exs"), g[t + "Exs"]) < TOL and relnorm(sim.history("phis"), g[t + "phis"]) < TOL


def _random_species(shape, NX, NY, Lx, Ly, charge, mass, seed, ppc=5, dt=0.01):
    rng = np.random.default_rng(seed)
    P = NX * NY * ppc
    vth = 0.3 * min(Lx / NX, Ly / NY) / dt  # a thermal particle crosses ~0.3 cell per step: cell changes and seam crossings occur
    return dict(x=Lx * (1 - rng.random(P)), y=Ly * (1 - rng.random(P)), vx=rng.standard_normal(P) * vth, vy=rng.standard_normal(P) * vth,
                vz=rng.standard_normal(P) * vth, charge=charge, mass=mass, weight=4 * math.pi ** 2 * Lx * Ly / P / abs(charge), shape=shape)


@pytest.mark.parametrize("shapes", [(0, 1), (10, 11), (12, 13), (14, 15), (15, 12)])
@pytest.mark.parametrize("acc", [1, 0])
@pytest.mark.parametrize("sort_every", [-1, 3])  # any-order kernel (global atomics) / tile-sorted kernel (shared-memory windows)
def test_loop_matches_oracle(es, oracle, shapes, acc, sort_every):
    NX, NY, Lx, Ly = 32, 16, 1.5, 2.0
    sp = [_random_species(shapes[0], NX, NY, Lx, Ly, -1.0, 1.0, 11), _random_species(shapes[1], NX, NY, Lx, Ly, 2.0, 7.0, 12)]
    sim, _ = _run_against(es, oracle, sp, NX, NY, Lx, Ly, 0.01, [0.9, -0.4, 0.6], NT=9, ntskip=2, ngskip=4, acc=acc, sort_every=sort_every)
    assert sim.sort_stats()[0] == (3 if sort_every > 0 else 0)  # sorted before steps 0, 3, 6


@pytest.mark.parametrize("kernel", ["tiled", "stream11"])
@pytest.mark.parametrize("shapes", [(0, 1), (12, 13), (15, 11), (14, 10)])
def test_tile_sorted_kernel_variants_match_oracle(es, oracle, monkeypatch, kernel, shapes):
    """The tile-sorted path has three particle kernels: es_particles_stream with replicated windows (the default, exercised by
    every sort_every > 0 test above), the same without replicas (PICGOLF_ES_KERNEL=stream11) and es_particles_tiled
    (work items, plain loads: PICGOLF_ES_KERNEL=tiled).  All shapes of the other two against the oracle as well."""
    monkeypatch.setenv("PICGOLF_ES_KERNEL", kernel)
    NX, NY, Lx, Ly = 32, 16, 1.5, 2.0
    sp = [_random_species(shapes[0], NX, NY, Lx, Ly, -1.0, 1.0, 21), _random_species(shapes[1], NX, NY, Lx, Ly, 2.0, 7.0, 22)]
    sim, _ = _run_against(es, oracle, sp, NX, NY, Lx, Ly, 0.01, [0.9, -0.4, 0.6], NT=7, ntskip=2, ngskip=4, acc=1, sort_every=2)
    assert sim.sort_stats()[0] == 4


@pytest.mark.parametrize("shape", [1, 12, 15])
def test_tiled_path_at_size(es, shape):
    """2^20 particles per species on a 64 x 128 grid (several tiles, several work items per tile): the tile-sorted kernel
    against the any-order kernel, particles returned in the caller's order, few window misses."""
    NX, NY, Lx, Ly, NT = 64, 128, 1.0, 2.0, 10
    P = 1 << 20
    n0 = 4 * math.pi ** 2
    dl = min(Lx / NX, Ly / NY)
    vth = dl * math.sqrt(n0)
    out = []
    for sort_every in (-1, 0):
        plasma = [es.Species(P, vth, n0, _shape(es, shape), Lx=Lx, Ly=Ly, charge=-1, mass=1),
                  es.Species(P, vth / 4, n0, _shape(es, shape), Lx=Lx, Ly=Ly, charge=1, mass=16)]
        sim = es.Simulation(plasma, es.ElectrostaticField(NX, NY, Lx, Ly, dt=dl / (6 * vth), B0x=math.sqrt(n0) / 4, accumulate=False),
                            es.ElectrostaticDiagnostics(NX, NY, NT, 2, 2), sort_every=sort_every)
        sim.loop(NT)
        out.append((sim.fields(), sim.species(0), sim.species(1), sim.scalars(), sim.history("Exs"), sim.sort_stats()))
    (fa, a0, a1, sa, ha, sta), (fb, b0, b1, sb, hb, stb) = out
    assert sta == (0, 0) and stb[0] == 1 and stb[1] < 1e-3 * 2 * P * NT  # auto = tiled at this size: sorted before step 0, the next sort is due after 16 steps
    assert relnorm(fb["rho"], fa["rho"]) < TOL_TILED and relnorm(fb["Ex"], fa["Ex"]) < TOL_TILED
    for x, y in list(zip(a0, b0)) + list(zip(a1, b1)):
        assert relnorm(y, x) < TOL_TILED  # same particle order as the caller's
    assert relnorm(sb["kineticenergy"], sa["kineticenergy"]) < TOL_TILED and relnorm(sb["fieldenergy"], sa["fieldenergy"]) < TOL_TILED
    assert relnorm(hb, ha) < TOL_TILED


def _species_n(shape, P, NX, NY, Lx, Ly, charge, mass, seed, dt):
    rng = np.random.default_rng(seed)
    vth = 0.3 * min(Lx / NX, Ly / NY) / dt
    return dict(x=Lx * (1 - rng.random(P)), y=Ly * (1 - rng.random(P)), vx=rng.standard_normal(P) * vth, vy=rng.standard_normal(P) * vth,
                vz=rng.standard_normal(P) * vth, charge=charge, mass=mass, weight=4 * math.pi ** 2 * Lx * Ly / max(P, 64) / abs(charge), shape=shape)


@pytest.mark.parametrize("sort_every", [-1, 2])
def test_smallest_grid_window_wraps_onto_itself(es, oracle, sort_every):
    """16 x 16 cells: one tile whose 32 x 32 shared-memory window covers the periodic grid four times over."""
    NX = NY = 16
    sp = [_species_n(15, 1500, NX, NY, 1.0, 1.0, -1.0, 1.0, 41, 0.02), _species_n(1, 700, NX, NY, 1.0, 1.0, 1.0, 3.0, 42, 0.02)]
    _run_against(es, oracle, sp, NX, NY, 1.0, 1.0, 0.02, [0.3, 0.2, -0.7], NT=6, ntskip=3, ngskip=1, acc=1, sort_every=sort_every)


@pytest.mark.parametrize("sort_every", [-1, 1])
def test_four_species_ragged_counts(es, oracle, sort_every):
    """The species limit, counts that are not multiples of anything (1, 31, 257, 1000), a different shape each."""
    NX, NY, Lx, Ly = 32, 64, 0.5, 3.0
    sp = [_species_n(sh, P, NX, NY, Lx, Ly, q, m, 50 + k, 0.01)
          for k, (sh, P, q, m) in enumerate(((12, 1, -1.0, 1.0), (0, 31, 1.0, 2.0), (13, 257, -2.0, 5.0), (14, 1000, 1.0, 11.0)))]
    _run_against(es, oracle, sp, NX, NY, Lx, Ly, 0.01, [0.0, 1.1, 0.4], NT=5, ntskip=1, ngskip=8, acc=0, sort_every=sort_every)


def test_three_species_and_reproducible_charge(es, oracle):
    NX = NY = 16
    sp = [_random_species(12, NX, NY, 1.0, 1.0, -1.0, 1.0, 21, dt=0.02), _random_species(13, NX, NY, 1.0, 1.0, 1.0, 4.0, 22, dt=0.02),
          _random_species(1, NX, NY, 1.0, 1.0, 1.0, 16.0, 23, dt=0.02)]
    for s in sp:
        s["weight"] /= 1.5
    sim, _ = _run_against(es, oracle, sp, NX, NY, 1.0, 1.0, 0.02, [0.0, 0.0, 1.0], NT=6, ntskip=1, ngskip=1, acc=0)
    sim2, _ = _run_against(es, oracle, sp, NX, NY, 1.0, 1.0, 0.02, [0.0, 0.0, 1.0], NT=6, ntskip=1, ngskip=1, acc=0)
    a, b = sim.fields(), sim2.fields()
    assert np.array_equal(a["rho"], b["rho"]) and np.array_equal(a["Ex"], b["Ex"])  # integer charge grid: order-free


def test_conservation_at_size(es):
    """Size-independent properties at 2 x 2^23 particles (128^2 grid, B-spline 2, tile-sorted by default at this size):
    total charge, momentum without a magnetic field, window misses."""
    NX = NY = 128
    P = 1 << 23
    n0 = 4 * math.pi ** 2
    dl = 1.0 / NX
    vth = dl * math.sqrt(n0)
    plasma = [es.Species(P, vth, n0, es.BSplineWeighting(2), Lx=1.0, Ly=1.0, charge=-1, mass=1),
              es.Species(P, vth / 4, 2 * n0, es.BSplineWeighting(2), Lx=1.0, Ly=1.0, charge=1, mass=16)]
    sim = es.Simulation(plasma, es.ElectrostaticField(NX, NY, 1.0, 1.0, dt=dl / (6 * vth), accumulate=False),
                        es.ElectrostaticDiagnostics(NX, NY, 12, 1, 4))
    sim.loop(12)
    rho = sim.fields()["rho"]
    dV = dl * dl
    q = [sp.charge * sp.weight * sp.P for sp in plasma]  # -n0 and +2 n0: net charge n0
    assert abs(rho.sum() * dV - sum(q)) < 1e-12 * sum(abs(v) for v in q)
    sc = sim.scalars()
    p, c = sc["particlemomentum"], sc["characteristicmomentum"]
    assert np.abs(p - p[0]).max() < 1e-11 * c.max()  # same shape for gather and deposit + spectral solve: momentum is conserved
    tot = sc["kineticenergy"] + sc["fieldenergy"]
    assert abs(tot[-1] / tot[0] - 1) < 0.05
    sorts, slow = sim.sort_stats()
    assert sorts == 1 and slow < 1e-4 * 2 * P * 12  # sorted before step 0; the next sort is due after 16 steps


def test_species_init_is_the_reference_halton_start(es, oracle):
    P, vth, n0, Lx, Ly = 4096, 0.013, 4 * math.pi ** 2, 2.0, 0.5
    sim = es.Simulation([es.Species(P, vth, n0, es.BSplineWeighting(2), Lx=Lx, Ly=Ly)], es.ElectrostaticField(16, 16, Lx, Ly, dt=0.01),
                        es.ElectrostaticDiagnostics(16, 16, 4, 1))
    x, y, vx, vy, vz = sim.species(0)
    xo, yo, vxo, vyo, vzo, w = oracle.es_species(P, vth, n0, Lx, Ly)
    assert np.array_equal(x, xo) and np.array_equal(y, yo)  # Halton positions: exact
    for a, b in ((vx, vxo), (vy, vyo), (vz, vzo)):
        assert relnorm(a, b) < 1e-12  # erfinv: CUDA vs scipy
        assert abs(a.mean()) < 1e-16 and a.std(ddof=1) == pytest.approx(vth / math.sqrt(2), rel=1e-13)
    assert sim.plasma[0].weight == w


def test_xyv_round_trip_and_resume(es):
    rng = np.random.default_rng(5)
    P = 3000
    a = np.column_stack([1 - rng.random(P), 1 - rng.random(P), rng.standard_normal((P, 3)) * 0.01])
    mk = lambda arr: es.Simulation([es.Species(P, 0.01, 1.0, es.AreaWeighting(), Lx=1.0, Ly=1.0, xyv=arr)],
                                   es.ElectrostaticField(16, 16, 1.0, 1.0, dt=0.05, B0x=1.0, accumulate=False), es.ElectrostaticDiagnostics(16, 16, 8, 1))
    sim = mk(a)
    assert np.array_equal(sim.xyv(0), a)
    sim2 = mk(a.T.copy())  # Julia's 5 x P layout
    assert np.array_equal(sim2.xyv(0), a)
    sim.loop(6)
    sim2.loop(3)
    mid, fl = sim2.xyv(0), sim2.fields()
    sim3 = mk(mid)
    sim3.set_field(fl["Exy_x"], fl["Exy_y"])
    sim3.loop(3)
    assert np.array_equal(sim3.xyv(0), sim.xyv(0))  # checkpoint / resume is bit-exact


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("mode", [0, 1])
def test_wk_spectrum_matches_numpy(es, axis, mode):
    rng = np.random.default_rng(9)
    NA, NB, ND = 32, 16, 64
    F = rng.standard_normal((NA, NB, ND))
    F += 3 * np.cos(2 * np.pi * (3 * np.arange(NA)[:, None, None] / NA + 5 * np.arange(ND)[None, None, :] / ND))
    got = es.wk_spectrum(F, axis=axis, mode=mode)
    if mode == 1:  # abs.(fft(F))[:, 1, :] resp. [1, :, :]      PIC2D3V.jl:1483,1489
        full = np.abs(np.fft.fftn(F))
        want = full[:, 0, :] if axis == 0 else full[0, :, :]
    elif axis == 0:  # sum(i -> abs.(fft(F[:, i, :])), 1:size(F, 2))      Electrostatic2D3V.jl:219
        want = sum(np.abs(np.fft.fft2(F[:, i, :])) for i in range(NB))
    else:            # sum(i -> abs.(fft(F[i, :, :])), 1:size(F, 1))      Electrostatic2D3V.jl:229
        want = sum(np.abs(np.fft.fft2(F[i, :, :])) for i in range(NA))
    assert got.shape == want.shape and relnorm(got, want) < TOL


def test_spectrum_of_the_stored_history(es):
    NX = NY = 16
    NT, ntskip = 32, 2
    sp = _random_species(12, NX, NY, 1.0, 1.0, -1.0, 1.0, 31, dt=0.02)
    a = np.stack([sp[k] for k in ("x", "y", "vx", "vy", "vz")], axis=1)
    s = es.Species(a.shape[0], 1.0, 1.0, es.BSplineWeighting(2), Lx=1.0, Ly=1.0, charge=-1.0, xyv=a)
    s.weight = sp["weight"]
    sim = es.Simulation([s], es.ElectrostaticField(NX, NY, 1.0, 1.0, dt=0.02, B0z=2.0, accumulate=False), es.ElectrostaticDiagnostics(NX, NY, NT, ntskip, 2))
    sim.loop(NT)
    for name in ("Exs", "Eys", "phis"):
        H = sim.history(name)
        assert H.shape == (8, 8, 16)
        full = np.abs(np.fft.fftn(H))
        scale = full.max()  # the k_y = 0 slice of Ey vanishes identically (Ey^ ~ k_y): compare on the scale of the whole transform
        assert np.abs(sim.spectrum(name, axis=0, mode=1) - full[:, 0, :]).max() < 1e-10 * scale
        assert np.abs(sim.spectrum(name, axis=1, mode=1) - full[0, :, :]).max() < 1e-10 * scale
        assert relnorm(sim.spectrum(name, axis=1, mode=0), sum(np.abs(np.fft.fft2(H[i])) for i in range(8))) < TOL
        assert relnorm(sim.spectrum(name, axis=0, mode=0), sum(np.abs(np.fft.fft2(H[:, i, :])) for i in range(8))) < TOL


def test_2d3v_snapshots_on_device(pg, es):
    """src/Electrostatic2D3V.jl:171-173: Exs, Eys, phis of every NS-th step kept on the device (field_history=1), and their
    omega-k map (lines 219-233) from the same library."""
    g = golden("c5_2d3v")
    NX, NY = int(g["NX"]), int(g["NY"])
    sim = pg.electrostatic_2d3v(NX=NX, NY=NY, P=int(g["P"]), T=8, NS=2, field_history=1)
    sim.set_particles(g["x0"], g["vx0"], y=g["y0"], vy=g["vy0"], vz=g["vz0"])
    sim.step(3)
    assert sim.snapshots("Exs").shape == (NX, NY, 1)  # t = 2 only
    sim.step(1)
    Exs, Eys, phis = sim.snapshots("Exs"), sim.snapshots("Eys"), sim.snapshots("phis")
    assert Exs.shape == (NX, NY, 2)
    for ti, t in enumerate((1, 3)):  # Julia t = 2, 4
        assert relnorm(Exs[:, :, ti].ravel(order="F"), g["Ex"][t]) < TOL and relnorm(Eys[:, :, ti].ravel(order="F"), g["Ey"][t]) < TOL
        rk = np.fft.fft2(g["rho"][t].reshape(NY, NX).T)
        rk[0, 0] = 0  # phi[1, 1] = 0   :144
        assert relnorm(phis[:, :, ti], np.real(np.fft.ifft2(rk))) < TOL
    K, _ = sim.diagnostics()
    assert K.shape == (2, 5) and relnorm(K[:, :3], g["K"][[1, 3], :3]) < TOL
    plain = pg.electrostatic_2d3v(NX=NX, NY=NY, P=int(g["P"]), T=8, NS=2)
    with pytest.raises(pg.PicGolfError):
        plain.snapshots("Exs")  # not kept unless asked for
    # the omega-k map of the snapshots: needs power-of-two extents -> repeat the two slices to 4
    F = np.concatenate([Exs, Exs], axis=2)
    want = sum(np.abs(np.fft.fft2(F[:, i, :])) for i in range(NY))  # sum(i -> abs.(fft(F[:, i, :])), 1:size(F, 2))   :219
    assert relnorm(es.wk_spectrum(F, axis=0, mode=0), want) < TOL


def test_argument_errors(es, pg):
    f = es.ElectrostaticField(16, 16, 1.0, 1.0, dt=0.01)
    d = es.ElectrostaticDiagnostics(16, 16, 4, 1)
    with pytest.raises(pg.PicGolfError) as e:
        es.Simulation([es.Species(10, 0.1, 1.0, es.AreaWeighting(), Lx=1, Ly=1)], es.ElectrostaticField(24, 16, dt=0.01), d)
    assert e.value.code == -5  # not a power of two: valid in the reference, not built
    sim = es.Simulation([es.Species(64, 0.1, 1.0, es.AreaWeighting(), Lx=1, Ly=1, xyv=np.full((64, 5), 0.5)),
                         es.Species(64, 0.1, 1.0, es.AreaWeighting(), Lx=1, Ly=1, xyv=np.full((64, 5), 0.5))], f, d)
    with pytest.raises(pg.PicGolfError):
        sim.set_species(2, *[np.zeros(64)] * 5)
    with pytest.raises(pg.PicGolfError):
        sim.set_species(0, *[np.zeros(63)] * 5)
    sim.loop(2)
    assert sim.steps_done == 2 and sim.launches >= 2 * 6
