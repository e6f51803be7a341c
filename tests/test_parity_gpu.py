"""Parity of the CUDA path (through the C ABI of libpicgolf.so) against the CPU oracle and the
committed golden fixtures.  Bars (BASELINE.json north_star):
  * NGP cell indices f(x), Julia mod, quiet start, NGP rho with dyadic w: BIT-EXACT;
  * rho, E, x, v after a step: |delta| <= 1e-12 * max|.| (norm-wise; SURVEY.md 7.4 explains why
    element-wise rtol is meaningless at zero crossings) -- at every size tested, N = 4096 and 256 x 256 included; the
    measured error of every comparison is written to PARITY.json (conftest.py), typically 1e-15 ... 5e-13;
  * diagnostics traces within 1e-9 relative over the short horizon before two-stream chaos diverges;
  * long horizon: growth rate of log10 D[:,1] within 1% of the analytic 2*gamma line."""
import math

import numpy as np
import pytest

from conftest import golden, relnorm

pytestmark = pytest.mark.gpu
TOL = 1e-12


# =============================================================================================
# stage level
# =============================================================================================
def test_ngp_index_bit_exact(pg, oracle):
    N = 128
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.random(100000), (np.arange(2 * N + 1) * 0.5) / N,  # every tie k+1/2 and integer
                        np.nextafter((np.arange(N) + 0.5) / N, 0), np.nextafter((np.arange(N) + 0.5) / N, 1),
                        [0.0, 1.0, 1e-300, 1 - 2 ** -53, -1e-20, 1.0000001]])
    for n in (64, 128, 4096):
        assert np.array_equal(pg.ngp_index(x, n), oracle.ngp_index(x, n))
    g = golden("c1_ngp")
    assert np.array_equal(pg.ngp_index(g["x0"], int(g["N"])), g["idx1"])


def test_float_mod_bit_exact(pg, oracle):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.random(10000) * 4 - 2, [-1e-20, -1e-300, 0.0, -0.0, 1.0, -1.0, 2.0, 1 - 2 ** -53,
                                                   -2 ** -54, 1.25, -0.25, 1e15 + 0.5]])
    got = pg.mod1(x)
    want = np.array([oracle.jl_mod1(v) for v in x])
    assert np.array_equal(got, want)
    assert got[10000] == 1.0  # mod(-1e-20, 1) == 1.0 exactly


def test_quiet_start_bit_exact(pg, oracle):
    for P, first, count in ((2048, 0, 2048), (2 ** 20, 12345, 4097), (2 ** 30, 2 ** 29 - 5, 1000)):
        x, v = pg.quiet_start(P, first, count)
        xo, vo = oracle.quiet_start(P, first, count)
        assert np.array_equal(x, xo) and np.array_equal(v, vo)
    x, _ = pg.quiet_start(2048)
    assert np.array_equal(x[:8], [0.5, 0.0, 0.75, 0.25, 0.625, 0.125, 0.875, 0.375])


def test_gauss_stencil_parity(pg, oracle):
    """Indices bit-exact; weights within 2.3e-16 absolute (2 ulp of 1/2 -- the reference's own erf/2
    quantisation) of the literal 2-erf evaluation, from the committed fixture."""
    g = golden("stencils")
    for N, hw in ((64, 7), (128, 6), (4096, 6), (4096, 7)):
        c = g[f"c_{N}_{hw}"]
        idx, wt = pg.gauss_stencil(c, N, hw)
        assert np.array_equal(idx, g[f"idx_{N}_{hw}"])
        assert np.abs(wt - g[f"wt_{N}_{hw}"]).max() < 2.3e-16
        assert np.abs(wt.sum(axis=1) - 1).max() < 5e-16
        if hw == 7:
            assert np.all(wt[:, 0] == 0) and np.all(wt[:, -1] == 0)


def test_ngp_deposit_bit_exact(pg, oracle):
    g = golden("c1_ngp")
    N, w = int(g["N"]), float(g["w"])
    rho = pg.ngp_deposit(g["x0"], N, w)
    assert np.array_equal(rho, oracle.ngp_deposit(g["x0"], N, w))  # dyadic w: exact and order independent
    rng = np.random.default_rng(2)
    x = rng.random(1 << 20)
    assert np.array_equal(pg.ngp_deposit(x, 4096, 0.5), oracle.ngp_deposit(x, 4096, 0.5))
    # ragged / degenerate inputs
    assert np.array_equal(pg.ngp_deposit(np.full(1000, 0.5), 128, 2.0), oracle.ngp_deposit(np.full(1000, 0.5), 128, 2.0))
    assert np.array_equal(pg.ngp_deposit(np.array([0.25]), 128, 1.0), oracle.ngp_deposit(np.array([0.25]), 128, 1.0))
    assert np.array_equal(pg.ngp_deposit(np.zeros(0), 128, 1.0), np.zeros(128))


@pytest.mark.parametrize("N,hw,P", [(128, 6, 4096), (64, 7, 2048), (4096, 6, 1 << 18), (4096, 7, 33333)])
def test_gauss_deposit_parity(pg, oracle, N, hw, P):
    rng = np.random.default_rng(3)
    x = rng.random(P) * 1.02 - 0.01  # slightly outside [0,1]: x is not wrapped during sweeps
    y = rng.random(P)
    w = 400.0 / P * N
    got = pg.gauss_deposit(x, y, N, hw, w)
    want = oracle.gauss_deposit(x, y, N, hw, w)
    assert relnorm(got, want) < TOL
    assert abs(got.sum() / (P * w) - 1) < 1e-13  # charge conservation
    # collisions: every particle in one cell
    xs = np.full(5000, 0.4321)
    assert relnorm(pg.gauss_deposit(xs, xs, N, hw, w), oracle.gauss_deposit(xs, xs, N, hw, w)) < TOL


def test_gauss_gather_parity(pg, oracle):
    rng = np.random.default_rng(4)
    for N, hw in ((128, 6), (64, 7), (4096, 6)):
        E = rng.standard_normal(N)
        c = np.concatenate([rng.random(2000), [-0.004, 1.003, 0.0, 1.0]])
        assert relnorm(pg.gauss_gather(E, c, hw), oracle.gauss_gather(E, c, N, hw)) < TOL


@pytest.mark.parametrize("N", [16, 64, 128, 1024, 4096, 8192])
def test_solve1d_parity(pg, oracle, N):
    rng = np.random.default_rng(5)
    rho = 200 + rng.standard_normal(N)
    assert relnorm(pg.solve1d(rho), oracle.solve1d(rho)) < TOL
    xs = np.arange(1, N + 1) / N
    E = pg.solve1d(200 + 0.7 * np.cos(2 * np.pi * 3 * xs))
    assert np.abs(E - 0.7 * np.sin(2 * np.pi * 3 * xs) / (2 * np.pi * 3)).max() < 1e-15 * N
    assert np.abs(pg.solve1d(200 + np.cos(2 * np.pi * (N // 2) * xs))).max() < 1e-13  # Nyquist -> 0
    assert np.abs(pg.solve1d(np.full(N, 5.0))).max() == 0.0                            # xi[1] *= 0
    # linearity (size-independent property)
    a, b = rng.standard_normal(N), rng.standard_normal(N)
    assert relnorm(pg.solve1d(a + 2 * b), pg.solve1d(a) + 2 * pg.solve1d(b)) < 1e-13


def test_unsupported_grid_is_an_error(pg):
    """Odd grids (the reference's own ik vector is malformed there) and erf-shape schemes on grids that are not 2^k."""
    with pytest.raises(pg.PicGolfError) as e:
        pg.solve1d(np.ones(101))
    assert e.value.code == -5
    with pytest.raises(pg.PicGolfError) as e:
        pg.gaussian_fixed_point(N=100, P=3200, T=4)
    assert e.value.code == -5
    with pytest.raises(pg.PicGolfError) as e:
        pg.ngp_fourier(N=101, P=6464, NT=4)
    assert e.value.code == -5


@pytest.mark.parametrize("N", [18, 100, 250, 1000, 6000])
def test_solve1d_any_even_grid(pg, oracle, N):
    """`fft` in NGPFourier.jl:3-5 takes any N: grids that are not a power of two go through direct transforms
    (solve1d_dft_fwd / solve1d_dft_inv) and must give the oracle's field (its O(N^2) long-double DFT)."""
    rng = np.random.default_rng(N)
    rho = rng.standard_normal(N)
    rho[0] += 3.0  # a DC component that xi[1] *= 0 must remove
    E = pg.solve1d(rho)
    Eo = oracle.solve1d(rho)
    assert relnorm(E, Eo) < TOL
    k = 3  # single mode in closed form: rho = cos(2 pi k x)  ->  E = sin(2 pi k x) / (2 pi k)
    x = np.arange(N) / N
    Em = pg.solve1d(np.cos(2 * np.pi * k * x))
    assert np.abs(Em - np.sin(2 * np.pi * k * x) / (2 * np.pi * k)).max() < 1e-13


@pytest.mark.parametrize("N,P", [(100, 6400), (96, 1 << 16), (1000, (1 << 21) + 7)])
def test_ngp_steps_on_grids_that_are_not_a_power_of_two(pg, oracle, N, P):
    """NGPFourier.jl:1-6 with N = 100, 96, 1000: the cell index goes through Julia's mod1 (no mask), the solve through the direct
    transforms; rho, E, x, v to 1e-12 over 6 steps, cell indices bit-exact.  The largest case takes
    the TMA-staged kernel of big shards."""
    rng = np.random.default_rng(N + 1)
    sim = pg.ngp_fourier(N=N, P=P, NT=8)
    x0 = rng.random(P)
    v0 = np.where(np.arange(1, P + 1) > P / 2, 1.0, -1.0)
    sim.set_particles(x0, v0)
    xo, vo = x0.copy(), v0.copy()
    for _ in range(6):
        sim.step(1)
        ro, Eo, _ = oracle.ngp_step(xo, vo, N, sim.cfg.dt, sim.cfg.w)
        rho, E = sim.fields()
        assert relnorm(rho, ro) < TOL
        assert relnorm(E, Eo) < TOL
    x, v = sim.particles()
    assert relnorm(x, xo) < TOL
    assert relnorm(v, vo) < TOL
    assert np.array_equal(pg.ngp_index(x0, N), oracle.ngp_index(x0, N))


def test_2d_stage_parity(pg, oracle):
    rng = np.random.default_rng(6)
    for NX, NY in ((32, 32), (64, 128), (256, 256)):
        rho = 39.0 + rng.standard_normal(NX * NY)
        Ex, Ey = pg.solve2d(rho, NX, NY)
        Exo, Eyo = oracle.solve2d(rho, NX, NY)
        assert relnorm(Ex, Exo) < TOL and relnorm(Ey, Eyo) < TOL
        P = 20000
        x, y = 1 - rng.random(P), 1 - rng.random(P)
        x[:4] = [1.0, 1e-12, 0.5, 1.0 / NX]  # grid points: r == 0 (the reference would @assert; weights (1,0))
        w = 39.0 / P * NX * NY
        assert relnorm(pg.cic_deposit(x, y, NX, NY, w), oracle.cic_deposit(x, y, NX, NY, w)) < TOL
        ex, ey = pg.cic_gather(Exo, Eyo, NX, NY, x, y)
        exo, eyo = oracle.cic_gather(Exo, Eyo, NX, NY, x, y)
        assert relnorm(ex, exo) < TOL and relnorm(ey, eyo) < TOL
    v = rng.standard_normal((3, 1000))
    E = rng.standard_normal((2, 1000))
    got = pg.boris(v[0], v[1], v[2], E[0], E[1], 0.013, 1.57)
    want = np.array([oracle.boris(v[0, i], v[1, i], v[2, i], E[0, i], E[1, i], 0.013, 1.57) for i in range(1000)]).T
    assert np.array_equal(np.array(got), want)  # same expression order -> bit-exact


# =============================================================================================
# full steps
# =============================================================================================
def test_c1_ngp_steps(pg, oracle):
    """Config 1 (src/NGPFourier.jl): rho bit-exact (dyadic w); E, x, v within 1e-12 per step."""
    g = golden("c1_ngp")
    N = int(g["N"])
    sim = pg.ngp_fourier(N=N, NT=64)
    assert sim.cfg.P == int(g["P"]) and sim.cfg.dt == float(g["dt"]) and sim.cfg.w == float(g["w"])
    sim.set_particles(g["x0"], g["v0"])
    x, v = g["x0"].copy(), g["v0"].copy()
    for t in range(8):
        sim.step(1)
        rho, E = sim.fields()
        ro, Eo, raw = oracle.ngp_step(x, v, N, sim.cfg.dt, sim.cfg.w)
        assert np.array_equal(rho, ro) and np.array_equal(rho, g["rho"][t])
        assert relnorm(E, Eo) < TOL and relnorm(E, g["E"][t]) < TOL
        xg, vg = sim.particles()
        assert relnorm(xg, x) < TOL and relnorm(vg, v) < TOL
        assert np.array_equal(pg.ngp_index(xg, N), oracle.ngp_index(x, N))
    R = sim.raw_diagnostics()
    assert R.shape == (8, 4) and relnorm(R[:, :3], g["raw"]) < TOL
    # stepping 8 at once (fused kick + next deposit passes) gives the same state as 8 single steps
    sim2 = pg.ngp_fourier(N=N, NT=64)
    sim2.set_particles(g["x0"], g["v0"])
    sim2.step(8)
    x2, v2 = sim2.particles()
    assert np.array_equal(x2, xg) and np.array_equal(v2, vg)
    assert relnorm(x2, g["x"]) < TOL and relnorm(v2, g["v"]) < TOL


@pytest.mark.parametrize("P", [1 << 21, 8192])
def test_leapfrog_split_calls_cost_no_extra_pass(pg, P):
    """A call of picgolf_step ends on full-step positions but leaves the NEXT step's charge deposited (the last pass deposits at the
    half-drifted positions without storing them); the next call redoes that half drift in its first pass.  3 + 1 + 4 steps must give
    the bits of 8 steps at once, and cost 8 particle passes + the one after set_particles, not 11 (TMA-staged and plain kernels)."""
    N = 4096 if P > 8192 else 128
    sims = []
    for split in ((8,), (3, 1, 4)):
        sim = pg.ngp_fourier(N=N, P=P, NT=16, W=256.0)
        sim.init_synthetic(seed=11)
        l0 = sim.launches
        for k in split:
            sim.step(k)
            x, v = sim.particles()  # a read-back between the calls must see full-step positions and must not disturb the pending deposit
            assert 0 <= x.min() and x.max() <= 1
        sims.append((sim.particles(), sim.fields(), sim.raw_diagnostics(), sim.launches - l0))
    (pa, fa, Ra, la), (pb, fb, Rb, lb) = sims
    assert all(np.array_equal(a, b) for a, b in zip(pa + fa, pb + fb)) and np.array_equal(Ra, Rb)
    assert la == lb == 1 + 3 * 8  # first deposit pass, then (solve, pass, step_end) per step


def test_explicit_gaussian_steps(pg, oracle):
    """src/Gaussian.jl: leapfrog with the erf shape."""
    g = golden("gauss_explicit")
    sim = pg.gaussian(NX=int(g["N"]), NT=16)
    assert sim.cfg.w == float(g["scale"])
    sim.set_particles(g["x0"], g["v0"])
    for t in range(8):
        sim.step(1)
        rho, E = sim.fields()
        assert relnorm(rho, g["rho"][t]) < TOL and relnorm(E, g["E"][t]) < TOL
    x, v = sim.particles()
    assert relnorm(x, g["x"]) < TOL and relnorm(v, g["v"]) < TOL
    assert relnorm(sim.raw_diagnostics()[:, :3], g["raw"]) < TOL


def test_c2_fixed_point_steps(pg, oracle):
    """Config 2 (src/GaussianFixedPoint.jl): sweeps/step == 4; rho, E, x, v within 1e-12 after one
    step; diagnostics D within 1e-10 over 16 steps."""
    g = golden("c2_fixedpoint")
    sim = pg.gaussian_fixed_point(T=64)
    assert (sim.cfg.N, sim.cfg.P) == (int(g["N"]), int(g["P"]))
    sim.set_particles(g["x0"], g["v0"])
    sim.step(1)
    rho, E = sim.fields()
    x, v = sim.particles()
    assert relnorm(rho, g["rho"][0]) < TOL and relnorm(E, g["E"][0]) < TOL
    assert relnorm(x, g["x1"]) < TOL and relnorm(v, g["v1"]) < TOL
    assert 0 <= x.min() and x.max() <= 1
    sim.step(15)
    D, sw = sim.diagnostics()
    assert D.shape == (16, 4) and np.array_equal(sw, g["sweeps"]) and np.all(sw == 4)
    assert relnorm(D[:, :3], g["D"][:, :3]) < TOL
    assert np.abs(D[:, 3] - g["D"][:, 3]).max() < 1e-13
    x, v = sim.particles()
    assert relnorm(x, g["x"]) < TOL and relnorm(v, g["v"]) < TOL  # 16 steps of error growth
    rho, E = sim.fields()
    assert relnorm(E, g["E"][15]) < TOL


def test_c2_against_live_oracle_other_seed(pg, oracle):
    rng = np.random.default_rng(99)
    N, P = 128, 4096
    x0, v0 = rng.random(P), rng.choice([-1.0, 1.0], P)
    sim = pg.gaussian_fixed_point(T=8)
    sim.set_particles(x0, v0)
    fp = oracle.FixedPoint(x0, v0, N, sim.cfg.dt, sim.cfg.W, hw=6, rtol=1e-8)
    for t in range(3):
        sim.step(1)
        D4, raw, s = fp.step()
        x, v = sim.particles()
        rho, E = sim.fields()
        assert relnorm(x, fp.x) < TOL and relnorm(v, fp.v) < TOL and relnorm(E, fp.E) < TOL
        assert relnorm(rho, fp.r) < TOL
    D, sw = sim.diagnostics()
    assert list(sw) == [4, 4, 4]


def test_c3_quiet_short_horizon(pg, oracle):
    """Config 3 (src/GaussianFixedPointQuiet.jl), first 16 steps.  E is pure round-off here (max|E| ~
    1e-16), so E is compared with atol = 1e-12*max|rho|/(2 pi); x, v to 1e-12."""
    g = golden("c3_quiet")
    sim = pg.gaussian_fixed_point_quiet(T=64)
    sim.init_quiet()
    x0, v0 = sim.particles()
    xo, vo = oracle.quiet_start(int(g["P"]))
    assert np.array_equal(x0, xo) and np.array_equal(v0, vo)
    sim.step(16)
    x, v = sim.particles()
    rho, E = sim.fields()
    assert relnorm(x, g["x16"]) < TOL and relnorm(v, g["v16"]) < TOL
    assert np.abs(E - g["E16"]).max() < 1e-12 * np.abs(rho).max() / (2 * np.pi)
    D, sw = sim.diagnostics()
    assert np.abs(D[:, 1] - g["D"][:16, 1]).max() < 1e-13 and np.abs(D[:, 2] - 1).max() < 1e-13
    assert np.abs(D[:, 3]).max() < 1e-15
    assert sw.min() >= 2 and sw.max() <= 10


def test_c3_growth_rate_long_horizon(pg, oracle):
    """The acceptance test of the reference (GaussianFixedPointQuiet.jl:16-20): over the full T=2^13
    run the field energy grows ~30 decades along the analytic 2*gamma line; momentum stays at
    round-off and the energy error bounded (README.md:46-47,76).  Compared with the oracle's trace."""
    g = golden("c3_quiet")
    T, dt, W = int(g["T"]), float(g["dt"]), float(g["W"])
    sim = pg.gaussian_fixed_point_quiet()
    sim.init_quiet()
    sim.step(T)
    D, sw = sim.diagnostics()
    assert D.shape == (T, 4)
    t = np.arange(1, T + 1) * dt
    sel = (t > 1) & (t < 8)
    slope = np.polyfit(t[sel], np.log10(D[sel, 0]), 1)[0]
    pred = oracle.growth_slope(W)
    assert abs(pred - 3.1509) < 1e-4
    assert abs(slope / pred - 1) < 0.01
    slope_o = np.polyfit(t[sel], np.log10(g["D"][sel, 0]), 1)[0]
    assert abs(slope - slope_o) < 0.02
    assert np.abs(D[:, 3]).max() < 1e-13
    assert np.abs(1 - D[:, 2]).max() < 1e-2 and np.abs(1 - D[: T // 5, 2]).max() < 1e-13
    assert abs(np.log10(D[:, 0]).max() - np.log10(g["D"][:, 0]).max()) < 0.05
    assert abs(np.log10(D[T // 8 - 1, 0]) - np.log10(g["D"][T // 8 - 1, 0])) < 1.0  # round-off seeded: same decade
    assert sw.min() >= 2 and sw.max() <= 10


def test_c5_2d3v_steps(pg, oracle):
    g = golden("c5_2d3v")
    NX, NY = int(g["NX"]), int(g["NY"])
    sim = pg.electrostatic_2d3v(NX=NX, NY=NY, P=int(g["P"]), T=16, NS=1)
    assert sim.cfg.dt == float(g["dt"]) and sim.cfg.w == float(g["w"]) and sim.cfg.B0 == float(g["B0"])
    sim.set_particles(g["x0"], g["vx0"], y=g["y0"], vy=g["vy0"], vz=g["vz0"])
    for t in range(4):
        sim.step(1)
        rho, Ex, Ey = sim.fields()
        assert relnorm(rho.reshape(-1, order="F"), g["rho"][t]) < TOL
        assert relnorm(Ex.reshape(-1, order="F"), g["Ex"][t]) < TOL and relnorm(Ey.reshape(-1, order="F"), g["Ey"][t]) < TOL
    x, y, vx, vy, vz = sim.particles()
    for a, k in ((x, "x"), (y, "y"), (vx, "vx"), (vy, "vy"), (vz, "vz")):
        assert relnorm(a, g[k]) < TOL
    assert x.min() > 0 and x.max() <= 1 and y.min() > 0 and y.max() <= 1
    K, _ = sim.diagnostics()
    assert K.shape == (4, 5) and relnorm(K[:, :3], g["K"][:, :3]) < TOL
    assert np.abs(K[:, 3:] - g["K"][:, 3:]).max() < 1e-14


# =============================================================================================
# size-independent properties at BASELINE sizes
# =============================================================================================
def test_scaled_config4_properties(pg, oracle):
    """Config 4 shape (N=4096, Gaussian fixed point) at 2^22 particles: the oracle checks a sampled
    deposit; conservation laws check the rest."""
    N, P = 4096, 1 << 22
    sim = pg.gaussian_fixed_point(N=N, P=P, T=8, W=400.0)
    sim.init_quiet()
    sim.step(2)
    x, v = sim.particles()
    rho, E = sim.fields()
    D, sw = sim.diagnostics()
    assert abs(rho.mean() / 400.0 - 1) < 1e-12           # charge conservation: mean rho == W
    assert abs(E.sum()) < 1e-9 * np.abs(E).max() * N + 1e-20  # xi[1]=0: zero-mean field
    assert np.abs(D[:, 3]).max() < 1e-14                 # momentum
    assert np.abs(D[:, 2] - 1).max() < 1e-12
    assert 0 <= x.min() and x.max() <= 1
    # seeded-uniform start, one step vs the oracle on the full population (seconds on CPU at 2^18)
    P2 = 1 << 18
    rng = np.random.default_rng(12)
    x0, v0 = rng.random(P2), np.where(np.arange(P2) >= P2 // 2, 1.0, -1.0)
    sim = pg.gaussian_fixed_point(N=N, P=P2, T=8, W=400.0)
    sim.set_particles(x0, v0)
    sim.step(1)
    fp = oracle.FixedPoint(x0, v0, N, sim.cfg.dt, 400.0, hw=6, rtol=1e-8)
    _, _, s = fp.step()
    x, v = sim.particles()
    rho, E = sim.fields()
    _, sw = sim.diagnostics()
    assert sw[0] == s
    assert relnorm(rho, fp.r) < TOL and relnorm(E, fp.E) < TOL
    assert relnorm(x, fp.x) < TOL and relnorm(v, fp.v) < TOL


def test_config4_full_size_properties(pg):
    """BASELINE.json config 4 at its full single-GPU size (N=4096, 2^28 particles, quiet two-stream start) through the
    path a caller gets (AUTO -> cell-polynomial passes): size-independent properties, and agreement of two different
    algorithms (polynomial passes vs the 13-weight window kernel) on the same 2^26-particle run."""
    N, P = 4096, 1 << 28
    sim = pg.gaussian_fixed_point(N=N, P=P, T=8, W=400.0, l=1e-8)
    assert sim.deposit_path == pg.DEPOSIT_POLY
    sim.init_quiet()
    sim.step(4)
    rho, E = sim.fields()
    D, sw = sim.diagnostics()
    assert abs(rho.mean() / 400.0 - 1) < 1e-12           # charge conservation: mean rho == W
    assert abs(E.sum()) < 1e-9 * np.abs(E).max() * N + 1e-20  # xi[1]=0: zero-mean field
    assert np.abs(D[:, 3]).max() < 1e-14                 # mean momentum of the symmetric beams
    assert np.abs(D[:, 2] - 1).max() < 1e-12             # energy: D3 = 1 while the field is round-off
    assert sw.max() <= 3
    x, v = sim.particles()
    assert 0 <= x.min() and x.max() <= 1 and np.abs(np.abs(v) - 1).max() < 1e-12  # round-off field: the beams stay cold
    xq, vq = pg.quiet_start(P, 12345, 1000)              # caller's order is restored: particle j drifted by v*t
    assert np.abs(np.mod(x[12345:13345] - xq - 4 * sim.cfg.dt * vq + 0.5, 1) - 0.5).max() < 1e-12
    sim.close(); del x, v
    P2 = 1 << 26
    out = []
    for mode in (pg.DEPOSIT_POLY, pg.DEPOSIT_SORTED):
        s2 = pg.gaussian_fixed_point(N=N, P=P2, T=8, W=400.0, l=1e-8, deposit_mode=mode)
        s2.init_synthetic(seed=3)
        s2.step(5)
        out.append(s2.fields() + s2.particles() + (s2.diagnostics()[1],))
        s2.close()
    (ra, Ea, xa, va, swa), (rb, Eb, xb, vb, swb) = out
    assert np.array_equal(swa, swb)
    assert relnorm(ra, rb) < TOL and relnorm(Ea, Eb) < 5e-12 and relnorm(xa, xb) < TOL and relnorm(va, vb) < TOL  # two algorithms, 5 steps, 2^26 particles: measured 5.7e-13


def test_scaled_ngp_properties(pg, oracle):
    """Config-1-scaled NGP (N=4096): rho is an exact histogram whatever the size."""
    N, P = 4096, 1 << 22
    sim = pg.ngp_fourier(N=N, P=P, NT=8, W=256.0)  # w = 256/2^22*4096 = 0.25 dyadic
    sim.init_synthetic(seed=7)
    x0, v0 = sim.particles()
    sim.step(1)
    rho, E = sim.fields()
    xh = x0.copy()
    oracle.ngp_half_drift(xh, v0, sim.cfg.dt)
    assert np.array_equal(rho, oracle.ngp_deposit(xh, N, sim.cfg.w))
    assert rho.sum() == P * sim.cfg.w
    x, v = sim.particles()
    ro, Eo, _ = oracle.ngp_step(x0, v0, N, sim.cfg.dt, sim.cfg.w)  # in place on x0, v0
    assert relnorm(E, Eo) < TOL and relnorm(x, x0) < TOL and relnorm(v, v0) < TOL


@pytest.mark.parametrize("mode", ["anyorder", "tiled"])
def test_c5_full_step_on_its_own_grid(pg, oracle, mode):
    """BASELINE.json config 5 on its own 256 x 256 grid (src/Electrostatic2D3V.jl:23 shape, 2^21 particles, Maxwellian start
    passed in): three full steps through the any-order kernel and through the tile-sorted kernel (first step any-order, sort,
    two tiled steps) against the multi-threaded oracle -- rho, Ex, Ey, the five particle arrays and K, at the 1e-12 bar."""
    NX = NY = 256
    P = 1 << 21
    rng = np.random.default_rng(55)
    sim = pg.electrostatic_2d3v(NX=NX, NY=NY, P=P, T=8, NS=1, deposit_mode=pg.DEPOSIT_ATOMIC if mode == "anyorder" else pg.DEPOSIT_SORTED,
                                sort_every=2)
    st = [1 - rng.random(P), 1 - rng.random(P)] + [rng.standard_normal(P) * sim.vth / math.sqrt(2) for _ in range(3)]
    sim.set_particles(st[0], st[2], y=st[1], vy=st[3], vz=st[4])
    so = [a.copy() for a in st]
    Ex, Ey = np.zeros(NX * NY), np.zeros(NX * NY)
    nthreads = oracle.max_threads()
    for t in range(3):
        sim.step(1)
        ro = oracle.step_2d3v(*so, NX, NY, sim.cfg.dt, sim.cfg.B0, sim.cfg.w, Ex, Ey, nthreads=nthreads)
        rho, ex, ey = sim.fields()
        assert relnorm(rho.reshape(-1, order="F"), ro) < TOL
        assert relnorm(ex.reshape(-1, order="F"), Ex) < TOL and relnorm(ey.reshape(-1, order="F"), Ey) < TOL
    got = sim.particles()  # caller's order, whatever the tile sort did
    for a, b in zip(got, so):
        assert relnorm(a, b) < TOL
    K, _ = sim.diagnostics()
    k1 = float(np.mean(Ex ** 2 + Ey ** 2))
    k2 = float(np.sum(so[2] ** 2 + so[3] ** 2) * sim.cfg.w)
    assert relnorm(K[2, :2], np.array([k1, k2])) < TOL
    if mode == "tiled":
        assert sim.sort_stats()[0] >= 1 and sim.sort_stats()[1] == 0  # sorted once, nothing left its window


def test_config5_full_size_properties(pg):
    """BASELINE.json config 5 at its full single-GPU size (256 x 256, 2^28 particles) through the path a caller gets
    (AUTO -> tile-sorted kernel): size-independent properties of Electrostatic2D3V.jl's loop."""
    NX = NY = 256
    P = 1 << 28
    sim = pg.electrostatic_2d3v(NX=NX, NY=NY, P=P, T=8, NS=1)
    sim.init_synthetic(seed=11, vth=sim.vth)
    sim.step(4)
    rho, Ex, Ey = sim.fields()
    K, _ = sim.diagnostics()
    n0 = 4 * math.pi ** 2
    assert abs(rho.mean() / n0 - 1) < 1e-12                       # charge conservation: mean(rho) == n0 (w = n0/P/(dx*dy))
    assert abs(Ex.sum()) < 1e-9 * np.abs(Ex).max() * NX * NY      # phi[1,1] = 0: zero-mean field
    assert abs(Ey.sum()) < 1e-9 * np.abs(Ey).max() * NX * NY
    assert abs(K[-1, 3] - K[0, 3]) < 1e-12 * sim.vth              # mean x momentum: CIC gather/deposit are adjoint (B0 || x rotates vy, vz only)
    assert abs(K[-1, 2] / K[0, 2] - 1) < 1e-3                     # total energy over 4 steps of a quiet thermal plasma
    assert sim.sort_stats()[0] >= 1 and sim.sort_stats()[1] < 1e-6 * P * 4
    x, y, vx, vy, vz = sim.particles()
    assert x.min() > 0 and x.max() <= 1 and y.min() > 0 and y.max() <= 1  # unimod keeps (0, 1]
    # K[:,2] = sum((vx^2+vy^2) w), K[:,4:5] = sum(v)/P of the state that came back in the caller's order
    assert abs(float(np.sum(vx ** 2 + vy ** 2)) * sim.cfg.w / K[-1, 1] - 1) < 1e-11
    assert abs(float(vx.sum()) / P - K[-1, 3]) < 1e-14
    del x, y, vx, vy, vz
    sim.close()


def test_ngp_full_size_properties(pg, oracle):
    """Config-1-scaled NGP at the benchmark's size (N=4096, 2^28 particles, TMA-staged pass): rho is an exact histogram of
    the half-drifted positions whatever the size (dyadic w), checked against numpy on the full population."""
    N, P = 4096, 1 << 28
    sim = pg.ngp_fourier(N=N, P=P, NT=8, W=256.0)  # w = 256 * 4096 / 2^28 = 2^-8
    sim.init_synthetic(seed=7)
    x0, v0 = sim.particles()
    sim.step(1)
    rho, E = sim.fields()
    xh = np.mod(x0 + v0 / 2 * sim.cfg.dt, 1)      # u(): first half drift (v = +-1: no negative zero / 1.0 corner here)
    cells = (np.rint(xh * N).astype(np.int64) - 1) % N  # f(x) = mod1(round(x*N), N), 0-based
    del xh
    counts = np.bincount(cells, minlength=N)
    del cells
    assert np.array_equal(rho, counts * sim.cfg.w) and rho.sum() == P * sim.cfg.w
    assert abs(E.sum()) < 1e-9 * np.abs(E).max() * N
    x, v = sim.particles()
    assert 0 <= x.min() and x.max() <= 1
    # second half drift + kick with the library's own E (NGPFourier.jl:6): v += E[f(x)]*dt at the full-step position
    assert np.array_equal(v, v0 + E[(np.rint(x * N).astype(np.int64) - 1) % N] * sim.cfg.dt)
    D, _ = sim.diagnostics()
    assert abs(D[0, 3]) < 1e-12                    # momentum of the symmetric beams after one kick
    sim.close()


# =============================================================================================
# cell-sorted deposit mode (pg_sort.cuh + fp_pass_sorted): same results, caller's particle order
# =============================================================================================
def test_sorted_mode_c2_golden(pg, oracle):
    g = golden("c2_fixedpoint")
    sim = pg.gaussian_fixed_point(T=64, deposit_mode=pg.DEPOSIT_SORTED, sort_every=3)
    sim.set_particles(g["x0"], g["v0"])
    sim.step(1)
    rho, E = sim.fields()
    x, v = sim.particles()  # unsorted back to the caller's order
    assert relnorm(rho, g["rho"][0]) < TOL and relnorm(E, g["E"][0]) < TOL
    assert relnorm(x, g["x1"]) < TOL and relnorm(v, g["v1"]) < TOL
    sim.step(15)
    D, sw = sim.diagnostics()
    assert np.array_equal(sw, g["sweeps"])
    assert relnorm(D[:, :3], g["D"][:, :3]) < TOL
    x, v = sim.particles()
    assert relnorm(x, g["x"]) < TOL and relnorm(v, g["v"]) < TOL
    sorts, slow = sim.sort_stats()
    assert sorts == 5  # lazy first sort: before steps 1,4,7,10,13 (the first step runs on the any-order kernel)


@pytest.mark.parametrize("start", ["uniform", "quiet"])
def test_sorted_mode_matches_atomic_mode(pg, oracle, start):
    """N=4096, 2^20 particles (256 per cell): AUTO picks the sorted path; 12 steps with re-sorts must
    agree with the order-agnostic atomic path to round-off, and with the oracle after one step."""
    N, P = 4096, 1 << 20
    rng = np.random.default_rng(21)
    sims = []
    for mode in (pg.DEPOSIT_ATOMIC, pg.DEPOSIT_AUTO):
        sim = pg.gaussian_fixed_point(N=N, P=P, T=16, W=400.0, deposit_mode=mode, sort_every=4)
        if start == "quiet":
            sim.init_quiet()
        else:
            x0 = rng.random(P) if not sims else x0  # noqa: F821
            v0 = np.where(np.arange(P) >= P // 2, 1.0, -1.0)
            sim.set_particles(x0, v0)
        sims.append(sim)
    a, s = sims
    x0, v0 = a.particles()
    xs0, vs0 = s.particles()
    assert np.array_equal(x0, xs0) and np.array_equal(v0, vs0)
    a.step(1); s.step(1)
    fp = oracle.FixedPoint(x0, v0, N, a.cfg.dt, 400.0, hw=6, rtol=1e-8)
    _, _, so = fp.step()
    for sim in (a, s):
        x, v = sim.particles()
        rho, E = sim.fields()
        assert relnorm(x, fp.x) < TOL and relnorm(v, fp.v) < TOL and relnorm(rho, fp.r) < TOL
        # quiet start: E is pure round-off, compare on the scale of rho
        assert np.abs(E - fp.E).max() < max(1e-11 * np.abs(fp.E).max(), 1e-12 * np.abs(fp.r).max() / (2 * np.pi))
        assert sim.diagnostics()[1][0] == so
    a.step(11); s.step(11)
    xa, va = a.particles()
    xs, vs = s.particles()
    assert relnorm(xs, xa) < TOL and relnorm(vs, va) < TOL
    Da, swa = a.diagnostics()
    Ds, sws = s.diagnostics()
    if start == "uniform":
        assert np.array_equal(swa, sws)
    assert relnorm(Ds[:, 1:3], Da[:, 1:3]) < TOL
    sorts, slow = s.sort_stats()
    assert sorts == 3 and slow < P // 100  # before steps 1,5,9; almost everything stays inside its window


# =============================================================================================
# cell-polynomial mode (pg_kernels_poly.cuh): moment deposit + per-cell gather polynomials
# =============================================================================================
def test_poly_mode_c2_golden(pg, oracle):
    """Golden config-2 fixture through the polynomial passes.  32 particles per cell in random order: every lane
    flushes its moment sets all the time and the gather window is restaged every row -- correctness only."""
    g = golden("c2_fixedpoint")
    sim = pg.gaussian_fixed_point(T=64, deposit_mode=pg.DEPOSIT_POLY, sort_every=3)
    sim.set_particles(g["x0"], g["v0"])
    sim.step(1)  # any-order kernel (lazy first sort)
    sim.step(1)  # first polynomial step
    sim2 = pg.gaussian_fixed_point(T=64, deposit_mode=pg.DEPOSIT_ATOMIC)
    sim2.set_particles(g["x0"], g["v0"])
    sim2.step(2)
    for a, b in zip(sim.particles() + sim.fields(), sim2.particles() + sim2.fields()):
        assert relnorm(a, b) < TOL
    sim.step(14)
    D, sw = sim.diagnostics()
    assert np.array_equal(sw, g["sweeps"])
    assert relnorm(D[:, :3], g["D"][:, :3]) < TOL
    x, v = sim.particles()
    assert relnorm(x, g["x"]) < TOL and relnorm(v, g["v"]) < TOL
    # the stand-alone sort before step 1 (0-based), then fused into the passes of steps 3, 6, 9, 12, 15: an order serves three steps
    assert sim.sort_stats()[0] == 6 and sim.fused_sorts == 5


@pytest.mark.parametrize("start", ["uniform", "quiet"])
@pytest.mark.parametrize("N,P", [(4096, 1 << 20), (256, (1 << 22) + 77)])
def test_poly_mode_matches_atomic_mode(pg, oracle, start, N, P):
    """Polynomial passes against the order-agnostic atomic path over 12 steps with re-sorts, and against the oracle
    after the first polynomial step; 256 and 16384 particles per cell, ragged tail included."""
    rng = np.random.default_rng(22)
    sims = []
    for mode in (pg.DEPOSIT_ATOMIC, pg.DEPOSIT_POLY):
        sim = pg.gaussian_fixed_point(N=N, P=P, T=16, W=400.0, deposit_mode=mode, sort_every=4)
        if start == "quiet":
            sim.init_quiet()
        else:
            x0 = rng.random(P) if not sims else x0  # noqa: F821
            v0 = np.where(np.arange(P) >= P // 2, 1.0, -1.0)
            sim.set_particles(x0, v0)
        sims.append(sim)
    a, s = sims
    x0, v0 = a.particles()
    a.step(2); s.step(2)  # step 1 runs on the any-order kernel in both; step 2 is the first polynomial one
    do_oracle = P <= (1 << 20)
    if do_oracle:
        fp = oracle.FixedPoint(x0, v0, N, a.cfg.dt, 400.0, hw=6, rtol=1e-8)
        so = [fp.step()[2] for _ in range(2)]
    xa, va = a.particles()
    xs, vs = s.particles()
    ra, Ea = a.fields()
    rs, Es = s.fields()
    assert relnorm(xs, xa) < TOL and relnorm(vs, va) < TOL and relnorm(rs, ra) < TOL
    escale = max(1e-11 * np.abs(Ea).max(), 1e-12 * np.abs(ra).max() / (2 * np.pi))  # quiet start: E is round-off
    assert np.abs(Es - Ea).max() < escale
    if do_oracle:
        assert relnorm(xs, fp.x) < TOL and relnorm(vs, fp.v) < TOL and relnorm(rs, fp.r) < TOL
        assert np.abs(Es - fp.E).max() < escale
        # quiet start: E is round-off and the count of the first polynomial step is decided by noise (its first solve
        # sees the charge the any-order kernel deposited, the second one the moment form of the same charge)
        assert list(s.diagnostics()[1][:2 if start == "uniform" else 1]) == so[:2 if start == "uniform" else 1]
    a.step(10); s.step(10)
    xa, va = a.particles()
    xs, vs = s.particles()
    assert relnorm(xs, xa) < TOL and relnorm(vs, va) < TOL
    Da, swa = a.diagnostics()
    Ds, sws = s.diagnostics()
    if start == "uniform":
        assert np.array_equal(swa, sws)
    assert relnorm(Ds[:, 1:3], Da[:, 1:3]) < TOL
    sorts, flushes = s.sort_stats()
    assert sorts == 3 and s.fused_sorts == 2  # before step 1 (0-based), then inside the passes of steps 4 and 8
    if P // N >= 4096:
        # (cell, sign v) bins drift as a whole: each lane of the warp that streams a bin flushes once per polynomial interval
        # (8 per cell and beam: 32 * 16 * N / P = 0.03 flushes per particle and pass here, plus one per warp range);
        # lanes alternating between intervals would flush on a large share of their particles
        assert flushes / (P * float(sws[1:].sum())) < 0.1


@pytest.mark.parametrize("det", [0, 1])
@pytest.mark.parametrize("max_sweeps", [10, 2, 1])
def test_fused_resort_matches_standalone_sort(pg, monkeypatch, max_sweeps, det):
    """The re-sort fused into the passes of a step (pg_kernels_poly.cuh: bins counted by pass k-1, slots reserved and written by
    the final pass k) against the same run with the stand-alone counting sort (PICGOLF_FUSED_SORT=0) and against the any-order
    kernels: warm beams, an order that serves ONE step (sort_every=1), ragged tail, and the short steps -- max_sweeps = 2 (the
    counting pass is pass 1, which starts from v = V) and max_sweeps = 1 (no counting pass: the final pass writes unpermuted to
    the same buffers).  Particles come back in the caller's order through the ids the final pass carries along."""
    N, P, steps = 256, (1 << 22) + 77, 9
    rng = np.random.default_rng(31)
    x0 = rng.random(P)
    v0 = np.where(np.arange(P) >= P // 2, 1.0, -1.0) + 0.3 * rng.standard_normal(P)
    runs = []
    for fused, mode in ((1, pg.DEPOSIT_POLY), (0, pg.DEPOSIT_POLY), (0, pg.DEPOSIT_ATOMIC)):
        monkeypatch.setenv("PICGOLF_FUSED_SORT", str(fused))
        sim = pg.gaussian_fixed_point(N=N, P=P, T=16, W=400.0, deposit_mode=mode, sort_every=1, max_sweeps=max_sweeps,
                                      deterministic=det if mode == pg.DEPOSIT_POLY else 0)
        sim.set_particles(x0, v0)
        sim.step(steps)
        runs.append((sim, sim.particles(), sim.fields(), sim.diagnostics()))
    (f, pf, ff, df), (s, ps, fs_, ds), (a, pa, fa, da) = runs
    assert f.deposit_path == pg.DEPOSIT_POLY and f.fused_sorts == steps - 2 and s.fused_sorts == 0  # step 0: any-order kernel, step 1: stand-alone sort
    assert f.sort_stats()[0] == steps - 1 and s.sort_stats()[0] == steps - 1
    for got, ref in ((pf, ps), (pf, pa)):
        for g, r in zip(got, ref):
            assert relnorm(g, r) < TOL
    for got, ref in ((ff, fs_), (ff, fa)):  # fields of a warm plasma after 9 steps of differently ordered sums (3e-12 measured)
        for g, r in zip(got, ref):
            assert relnorm(g, r) < 2e-11
    assert np.array_equal(df[1], da[1]) and np.array_equal(ds[1], da[1])
    assert relnorm(df[0][:steps, 1:3], da[0][:steps, 1:3]) < TOL
    if max_sweeps == 1:
        assert set(df[1][:steps]) == {1}
    # launch accounting (sweeps are counted on the device): a re-sorting step launches the bin scan once per sweep, the other run
    # launched the three kernels of the stand-alone sort before each of those steps instead
    assert f.launches - s.launches == int(df[1][2:steps].sum()) - 3 * (steps - 2)
    # the fused order is the order of the next step's mid-points: a lane changes interval when the stream does, not because of the spread in v
    if not det and max_sweeps == 10:
        assert f.sort_stats()[1] < 0.1 * P * float(df[1][:steps].sum())
    for sim in (f, s, a):
        sim.close()


def test_fused_resort_reproducible_in_deterministic_mode(pg):
    """deterministic = 1 with a fused re-sort every step: the order inside a bin changes from run to run (atomic slot
    reservation), the integer moment sums do not care -- bit-identical rho, E, x, v, D."""
    N, P = 256, 1 << 22
    out = []
    for _ in range(2):
        sim = pg.gaussian_fixed_point(N=N, P=P, T=16, W=400.0, deposit_mode=pg.DEPOSIT_POLY, sort_every=1, deterministic=1)
        sim.init_synthetic(seed=5, vth=0.2)
        sim.step(8)
        assert sim.fused_sorts == 6
        out.append(sim.particles() + sim.fields() + (sim.diagnostics()[0][:8],))
        sim.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)


def test_poly_mode_warm_beams_force_resort(pg):
    """Warm beams shear a (cell, sign v) bin over several cells within a few steps: lanes start to alternate between
    cells, the per-step flush probe must force re-sorts long before the default interval, and the results must still
    agree with the any-order path."""
    N, P = 256, 1 << 22
    rng = np.random.default_rng(9)
    x0 = rng.random(P)
    v0 = np.where(np.arange(P) >= P // 2, 1.0, -1.0) + 0.6 * rng.standard_normal(P)
    a = pg.gaussian_fixed_point(N=N, P=P, T=32, W=400.0, deposit_mode=pg.DEPOSIT_ATOMIC)
    s = pg.gaussian_fixed_point(N=N, P=P, T=32, W=400.0, deposit_mode=pg.DEPOSIT_POLY)  # sort_every=0: adaptive
    for sim in (a, s):
        sim.set_particles(x0, v0)
        sim.step(24)
    xa, va = a.particles()
    xs, vs = s.particles()
    assert relnorm(xs, xa) < TOL and relnorm(vs, va) < TOL
    Da, swa = a.diagnostics()
    Ds, sws = s.diagnostics()
    assert np.array_equal(swa, sws) and relnorm(Ds[:, 1:3], Da[:, 1:3]) < TOL
    sorts, flushes = s.sort_stats()
    assert sorts >= 3  # the starting interval alone (16 steps) would give 2


def test_poly_mode_charge_and_reproducibility(pg):
    """Total charge is conserved to round-off by the moment form; two runs agree to round-off (the counting sort
    ranks particles inside a bin with atomics, so the order of the moment sums differs from run to run)."""
    N, P = 1024, 1 << 22
    out = []
    for _ in range(2):
        sim = pg.gaussian_fixed_point(N=N, P=P, T=8, W=400.0, deposit_mode=pg.DEPOSIT_POLY, sort_every=100)
        sim.init_synthetic(seed=5)
        sim.step(4)
        rho, E = sim.fields()
        out.append((rho, E) + sim.particles())
        assert abs(rho.mean() / 400.0 - 1) < 1e-13
    assert all(relnorm(a, b) < 1e-13 for a, b in zip(*out))


def test_2d3v_tile_sorted_mode(pg, oracle):
    """Tile-sorted 2D path (pg_sort mode 1 + particles_2d3v_tiled): golden fixture with frequent re-sorts,
    then a 128x128 run against the any-order path."""
    g = golden("c5_2d3v")
    NX, NY = int(g["NX"]), int(g["NY"])
    sim = pg.electrostatic_2d3v(NX=NX, NY=NY, P=int(g["P"]), T=16, NS=1, deposit_mode=pg.DEPOSIT_SORTED, sort_every=2)
    sim.set_particles(g["x0"], g["vx0"], y=g["y0"], vy=g["vy0"], vz=g["vz0"])
    for t in range(4):
        sim.step(1)
        rho, Ex, Ey = sim.fields()
        assert relnorm(rho.reshape(-1, order="F"), g["rho"][t]) < TOL
        assert relnorm(Ex.reshape(-1, order="F"), g["Ex"][t]) < TOL
    got = sim.particles()  # caller's order
    for a, k in zip(got, ("x", "y", "vx", "vy", "vz")):
        assert relnorm(a, g[k]) < TOL
    K, _ = sim.diagnostics()
    assert relnorm(K[:, :3], g["K"][:, :3]) < TOL
    assert sim.sort_stats()[0] == 2
    # larger: sorted (AUTO) vs any-order
    NX = NY = 128
    P = 1 << 20
    rng = np.random.default_rng(31)
    sims = [pg.electrostatic_2d3v(NX=NX, NY=NY, P=P, T=16, NS=1, deposit_mode=m, sort_every=5)
            for m in (pg.DEPOSIT_ATOMIC, pg.DEPOSIT_AUTO)]
    st = [1 - rng.random(P), 1 - rng.random(P)] + [rng.standard_normal(P) * sims[0].vth / np.sqrt(2) for _ in range(3)]
    for s in sims:
        s.set_particles(st[0], st[2], y=st[1], vy=st[3], vz=st[4])
        s.step(12)
    pa, ps = sims[0].particles(), sims[1].particles()
    for a, b in zip(ps, pa):
        assert relnorm(a, b) < TOL
    fa, fs = sims[0].fields(), sims[1].fields()
    assert relnorm(fs[0], fa[0]) < TOL  # both fixed point; the tiled path quantises per work item, not per deposit
    assert relnorm(fs[1], fa[1]) < TOL
    sorts, slow = sims[1].sort_stats()
    assert sorts == 3 and slow < P // 1000
    # one step vs the oracle
    so = [a.copy() for a in st]
    Ex, Ey = np.zeros(NX * NY), np.zeros(NX * NY)
    sim = pg.electrostatic_2d3v(NX=NX, NY=NY, P=P, T=4, NS=1)
    sim.set_particles(st[0], st[2], y=st[1], vy=st[3], vz=st[4])
    sim.step(1)
    ro = oracle.step_2d3v(*so, NX, NY, sim.cfg.dt, sim.cfg.B0, sim.cfg.w, Ex, Ey, nthreads=4)
    rho, ex, ey = sim.fields()
    assert relnorm(rho.reshape(-1, order="F"), ro) < TOL and relnorm(ex.reshape(-1, order="F"), Ex) < TOL
    got = sim.particles()
    assert relnorm(got[0], so[0]) < TOL and relnorm(got[2], so[2]) < TOL


def test_ngp_odd_particle_count(pg, oracle):
    """The NGP pass streams particle pairs with 128-bit accesses; an odd P exercises the tail."""
    N, P = 128, 8191
    rng = np.random.default_rng(41)
    x0 = rng.random(P)
    v0 = np.where(np.arange(P) >= P // 2, 1.0, -1.0)
    sim = pg.ngp_fourier(N=N, P=P, NT=8, W=200.0)
    sim.set_particles(x0, v0)
    sim.step(3)
    x, v = x0.copy(), v0.copy()
    for _ in range(3):
        ro, Eo, _ = oracle.ngp_step(x, v, N, sim.cfg.dt, sim.cfg.w)
    rho, E = sim.fields()
    xg, vg = sim.particles()
    assert relnorm(rho, ro) < 1e-13 and relnorm(E, Eo) < TOL and relnorm(xg, x) < TOL and relnorm(vg, v) < TOL


# =============================================================================================
# edge cases: ragged sizes, extreme grids, sweep cap, checkpoint/resume, trace capacity
# =============================================================================================
@pytest.mark.parametrize("P", [1, 33, 1000])
def test_ragged_particle_counts_atomic(pg, oracle, P):
    N = 64
    rng = np.random.default_rng(P)
    x0, v0 = rng.random(P), rng.choice([-1.0, 1.0], P)
    sim = pg.gaussian_fixed_point(N=N, P=P, T=8, W=50.0)
    sim.set_particles(x0, v0)
    sim.step(3)
    fp = oracle.FixedPoint(x0, v0, N, sim.cfg.dt, 50.0, hw=6, rtol=1e-8)
    sw_o = [fp.step()[2] for _ in range(3)]
    x, v = sim.particles()
    rho, E = sim.fields()
    D, sw = sim.diagnostics()
    assert list(sw) == sw_o
    assert relnorm(x, fp.x) < TOL and relnorm(v, fp.v) < TOL and relnorm(rho, fp.r) < TOL and relnorm(E, fp.E) < TOL


def test_ragged_particle_count_sorted(pg, oracle):
    """Sorted path with a particle count that is not a multiple of the warp chunk (tail groups, dead lanes)."""
    N, P = 1024, (1 << 18) + 37
    rng = np.random.default_rng(77)
    x0, v0 = rng.random(P), np.where(np.arange(P) % 2 == 0, 1.0, -1.0)
    sim = pg.gaussian_fixed_point(N=N, P=P, T=8, W=100.0, deposit_mode=pg.DEPOSIT_SORTED)
    sim.set_particles(x0, v0)
    sim.step(2)
    fp = oracle.FixedPoint(x0, v0, N, sim.cfg.dt, 100.0, hw=6, rtol=1e-8)
    sw_o = [fp.step()[2] for _ in range(2)]
    x, v = sim.particles()
    rho, E = sim.fields()
    assert list(sim.diagnostics()[1]) == sw_o
    assert relnorm(x, fp.x) < TOL and relnorm(v, fp.v) < TOL and relnorm(rho, fp.r) < TOL and relnorm(E, fp.E) < TOL


@pytest.mark.parametrize("N", [16, 8192])
def test_extreme_grid_sizes(pg, oracle, N):
    P = 4 * N if N == 8192 else 4096
    rng = np.random.default_rng(N)
    x0, v0 = rng.random(P), rng.choice([-1.0, 1.0], P)
    sim = pg.gaussian_fixed_point(N=N, P=P, T=4, W=20.0)
    sim.set_particles(x0, v0)
    sim.step(1)
    fp = oracle.FixedPoint(x0, v0, N, sim.cfg.dt, 20.0, hw=6, rtol=1e-8)
    _, _, s = fp.step()
    x, v = sim.particles()
    rho, E = sim.fields()
    assert sim.diagnostics()[1][0] == s
    assert relnorm(rho, fp.r) < TOL and relnorm(E, fp.E) < TOL and relnorm(x, fp.x) < TOL and relnorm(v, fp.v) < TOL
    ng = pg.ngp_fourier(N=N, P=P, NT=4, W=float(N))  # w = N*N/P dyadic
    ng.set_particles(x0, v0)
    ng.step(2)
    xo, vo = x0.copy(), v0.copy()
    for _ in range(2):
        ro, Eo, _ = oracle.ngp_step(xo, vo, N, ng.cfg.dt, ng.cfg.w)
    rho, E = ng.fields()
    xg, vg = ng.particles()
    assert np.array_equal(rho, ro) and relnorm(E, Eo) < TOL and relnorm(xg, xo) < TOL and relnorm(vg, vo) < TOL


def test_sweep_cap_is_not_an_error(pg, oracle):
    """rtol=0, atol=0 never converges on noisy data: the reference silently proceeds after 10 sweeps."""
    N, P = 128, 4096
    rng = np.random.default_rng(5)
    x0, v0 = rng.random(P), rng.choice([-1.0, 1.0], P)
    sim = pg.gaussian_fixed_point(T=4, l=0.0)
    sim.set_particles(x0, v0)
    sim.step(2)
    fp = oracle.FixedPoint(x0, v0, N, sim.cfg.dt, 400.0, hw=6, rtol=0.0)
    sw_o = [fp.step()[2] for _ in range(2)]
    D, sw = sim.diagnostics()
    x, v = sim.particles()
    assert list(sw) == sw_o
    assert sw.max() <= 10
    assert relnorm(x, fp.x) < TOL and relnorm(v, fp.v) < TOL
    few = pg.gaussian_fixed_point(T=4, max_sweeps=2)
    few.set_particles(x0, v0)
    few.step(1)
    assert few.diagnostics()[1][0] == 2


def test_checkpoint_resume_bit_exact(pg):
    """State is (x, v, E): get -> new handle -> set_particles + set_field continues bit-identically
    (order-free fixed-point deposits)."""
    rng = np.random.default_rng(8)
    x0, v0 = rng.random(4096), rng.choice([-1.0, 1.0], 4096)
    a = pg.gaussian_fixed_point(T=16)
    a.set_particles(x0, v0)
    a.step(6)
    xa, va = a.particles()
    b = pg.gaussian_fixed_point(T=16)
    b.set_particles(x0, v0)
    b.step(3)
    xm, vm = b.particles()
    _, Em = b.fields()
    c = pg.gaussian_fixed_point(T=16)
    c.set_particles(xm, vm)
    c.set_field(Em)
    c.step(3)
    xc, vc = c.particles()
    assert np.array_equal(xc, xa) and np.array_equal(vc, va)
    assert np.array_equal(c.fields()[1], a.fields()[1])
    assert c.steps_done == 3 and a.steps_done == 6


def test_trace_capacity_and_diag_every(pg):
    sim = pg.gaussian_fixed_point(T=4)
    sim.init_synthetic(seed=3)
    sim.step(7)  # more steps than rows: recording stops, stepping does not
    D, sw = sim.diagnostics()
    assert D.shape == (4, 4) and sim.steps_done == 7
    s2 = pg.electrostatic_2d3v(NX=32, NY=32, P=8192, T=8, NS=2)
    s2.init_synthetic(seed=4, vth=s2.vth)
    s2.step(7)
    K, _ = s2.diagnostics()
    assert K.shape == (3, 5)  # rows at t = 2, 4, 6  (Electrostatic2D3V.jl:164 `if t % NS == 0`)
    assert np.all(K[:, 1] > 0) and np.all(np.isfinite(K))


def test_deterministic_flag_bit_reproducible(pg):
    """deterministic=1: charge does not depend on particle order or on the run."""
    N, P = 4096, 1 << 19
    rng = np.random.default_rng(9)
    x0 = rng.random(P)
    v0 = np.where(np.arange(P) >= P // 2, 1.0, -1.0)
    perm = rng.permutation(P)
    res = []
    for xs, vs in ((x0, v0), (x0[perm], v0[perm]), (x0, v0)):
        sim = pg.gaussian_fixed_point(N=N, P=P, T=4, W=400.0, deterministic=1)
        sim.set_particles(xs, vs)
        sim.step(2)
        res.append((sim.fields(), sim.particles()))
    assert np.array_equal(res[0][0][0], res[1][0][0]) and np.array_equal(res[0][0][1], res[1][0][1])  # order free
    assert np.array_equal(res[0][1][0], res[2][1][0]) and np.array_equal(res[0][1][1], res[2][1][1])  # run to run
    assert np.array_equal(res[1][1][0], res[0][1][0][perm])


@pytest.mark.parametrize("N,P", [(1024, 1 << 22), (256, (1 << 22) + 77)])
def test_deterministic_fast_path_bit_reproducible(pg, N, P):
    """deterministic=1 at a size where AUTO picks the polynomial passes: the handle STAYS on fp_pass_poly (lanes sum integers,
    pg_kernels_poly.cuh DET) and rho, E, x, v and the diagnostics are bit-identical from run to run and for any order in which
    the caller hands the particles over -- although the counting sort ranks the particles of a bin with atomics, so the lanes
    sum different particles every run.  Against the default (fp64 lane sums) path the results agree to round-off."""
    rng = np.random.default_rng(19)
    x0 = rng.random(P)
    v0 = np.where(np.arange(P) >= P // 2, 1.0, -1.0) + 1e-3 * rng.standard_normal(P)
    perm = rng.permutation(P)
    res = []
    for xs, vs, det in ((x0, v0, 1), (x0, v0, 1), (x0[perm], v0[perm], 1), (x0, v0, 0)):
        sim = pg.gaussian_fixed_point(N=N, P=P, T=8, W=400.0, deterministic=det, sort_every=3)
        assert sim.deposit_path == pg.DEPOSIT_POLY
        sim.set_particles(xs, vs)
        sim.step(7)  # step 1 any-order, stand-alone sort before step 2, re-sorts fused into the passes of steps 4 and 7, six polynomial steps
        res.append(sim.fields() + sim.particles() + (sim.raw_diagnostics(), sim.diagnostics()[1]))
        assert sim.sort_stats()[0] == 3 and sim.fused_sorts == 2
        sim.close()
    a, b, c, d = res
    for u, w in zip(a, b):
        assert np.array_equal(u, w)                       # run to run
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])  # rho, E: any particle order
    assert np.array_equal(c[2], a[2][perm]) and np.array_equal(c[3], a[3][perm])
    # sum E^2, sum v^2, sum v, sweeps of the polynomial steps (row 0 is the any-order first step, whose fp64 block sums of v follow the caller's order)
    assert np.array_equal(a[4][1:], c[4][1:]) and np.array_equal(a[4][0, 0], c[4][0, 0]) and np.array_equal(a[5], c[5])
    assert relnorm(a[0], d[0]) < TOL and relnorm(a[2], d[2]) < TOL and relnorm(a[3], d[3]) < TOL  # vs the default path
    assert relnorm(a[4][:, :2], d[4][:, :2]) < TOL and np.array_equal(a[5], d[5])


# =============================================================================================
# Simpson-1/3 fixed point (src/GaussianFixedPointQuietSimpson13.jl)
# =============================================================================================
def test_simpson13_steps(pg, oracle):
    g = golden("simpson13")
    # noisy start, N=128, +-6, l=1e-8: sweeps equal, state within 1e-11 after 8 steps
    sim = pg.gaussian_fixed_point_quiet_simpson13(N=128, P=4096, T=16, W=400.0, l=1e-8, half_width=6)
    sim.set_particles(g["xr"], g["vr"])
    sim.step(8)
    D, sw = sim.diagnostics()
    x, v = sim.particles()
    rho, E = sim.fields()  # rho(x,x) and E[end,:]
    assert np.array_equal(sw, g["swn"])
    assert relnorm(D[:, :3], g["Dn"][:, :3]) < TOL and np.abs(D[:, 3] - g["Dn"][:, 3]).max() < 1e-13
    assert relnorm(x, g["xn"]) < TOL and relnorm(v, g["vn"]) < TOL
    assert relnorm(rho, g["rn"]) < TOL and relnorm(E, g["En"][2 * 128:]) < TOL
    # one step against the live oracle at 1e-12
    one = pg.gaussian_fixed_point_quiet_simpson13(N=128, P=4096, T=4, W=400.0, l=1e-8, half_width=6)
    one.set_particles(g["xr"], g["vr"])
    one.step(1)
    o1 = oracle.Simpson13(g["xr"], g["vr"], 128, one.cfg.dt, 400.0, hw=6, rtol=1e-8)
    _, _, s = o1.step()
    x, v = one.particles()
    rho, E = one.fields()
    assert one.diagnostics()[1][0] == s
    assert relnorm(x, o1.x) < TOL and relnorm(v, o1.v) < TOL and relnorm(rho, o1.r) < TOL and relnorm(E, o1.E[256:]) < TOL


def test_simpson13_quiet_growth(pg, oracle):
    """The script's own configuration (N=64, P=2048, +-7, l=4eps, quiet start) over its first 2048 steps:
    growth rate on the analytic line, energy conserved to round-off, state after 16 steps equal to the oracle's."""
    g = golden("simpson13")
    T, dt, W = int(g["T"]), float(g["dt"]), float(g["W"])
    sim = pg.gaussian_fixed_point_quiet_simpson13(T=T)
    assert (sim.cfg.N, sim.cfg.P, sim.cfg.half_width) == (64, 2048, 7) and sim.cfg.rtol == 4 * np.finfo(float).eps
    sim.init_quiet()
    sim.step(16)
    x, v = sim.particles()
    assert relnorm(x, g["x16"]) < TOL and relnorm(v, g["v16"]) < TOL
    sim.step(T - 16)
    D, sw = sim.diagnostics()
    t = np.arange(1, T + 1) * dt
    sel = (t > 1) & (t < 5)
    slope = np.polyfit(t[sel], np.log10(D[sel, 0]), 1)[0]
    assert abs(slope / oracle.growth_slope(W) - 1) < 0.01
    assert abs(slope - np.polyfit(t[sel], np.log10(g["D"][sel, 0]), 1)[0]) < 0.02
    assert np.abs(D[:, 3]).max() < 1e-13 and np.abs(1 - D[:, 2]).max() < 1e-12
    assert sw.min() >= 2 and sw.max() <= 10


def test_area_simpson13(pg, oracle):
    """src/AreaFixedPointQuietSimpson13.jl: Simpson-1/3 schedule with the 2-cell area shape d(y) (line 5), l=1e-14."""
    g = golden("area_simpson13")
    sim = pg.area_fixed_point_quiet_simpson13(N=128, P=4096, T=16, W=400.0, l=1e-9)
    sim.set_particles(g["xr"], g["vr"])
    sim.step(8)
    D, sw = sim.diagnostics()
    x, v = sim.particles()
    rho, E = sim.fields()
    assert np.array_equal(sw, g["swn"])
    assert relnorm(D[:, :3], g["Dn"][:, :3]) < TOL
    assert relnorm(x, g["xn"]) < TOL and relnorm(v, g["vn"]) < TOL and relnorm(rho, g["rn"]) < TOL
    assert relnorm(E, g["En"][256:]) < TOL
    T, dt, W = int(g["T"]), float(g["dt"]), float(g["W"])
    q = pg.area_fixed_point_quiet_simpson13(T=T)
    assert q.cfg.rtol == 1e-14 and (q.cfg.N, q.cfg.P) == (64, 2048)
    q.init_quiet()
    q.step(16)
    x, v = q.particles()
    assert relnorm(x, g["x16"]) < TOL and relnorm(v, g["v16"]) < TOL
    q.step(T - 16)
    D, sw = q.diagnostics()
    t = np.arange(1, T + 1) * dt
    sel = (t > 1) & (t < 2.6)
    slope = np.polyfit(t[sel], np.log10(D[sel, 0]), 1)[0]
    assert abs(slope / oracle.growth_slope(W) - 1) < 0.01
    assert np.abs(D[:, 3]).max() < 1e-13 and np.abs(1 - D[:, 2]).max() < 1e-12


# =============================================================================================
# 1D2V magnetised Boris code (src/NGP1D2V.jl)
# =============================================================================================
def test_ngp1d2v_steps(pg, oracle):
    g = golden("ngp1d2v")
    N = int(g["N"])
    sim = pg.ngp_1d2v(T=64, TO=16)  # window T/TO = 4 steps
    assert (sim.cfg.N, sim.cfg.P, sim.cfg.diag_every, sim.cfg.half_width) == (512, 7680, 4, 7)
    assert sim.cfg.dt == float(g["dt"]) and sim.cfg.B0 == float(g["B0"]) and sim.cfg.w == float(g["w"])
    sim.set_particles(g["x0"], g["vx0"], vy=g["vy0"])
    for t in range(8):
        sim.step(1)
        rho, E = sim.fields()
        assert relnorm(rho, g["rho"][t]) < TOL and relnorm(E, g["E"][t]) < TOL
    x, vx, vy = sim.particles()
    assert relnorm(x, g["x"]) < TOL and relnorm(vx, g["vx"]) < TOL and relnorm(vy, g["vy"]) < TOL
    assert 0 <= x.min() and x.max() <= 1
    # diagnostics: one row per window, formed as NGP1D2V.jl:59-61,64 from the sums at the window's last step
    D, _ = sim.diagnostics()
    assert D.shape == (2, 5)
    n0, P = float(g["n0"]), int(g["P"])
    for ti, t in enumerate((3, 7)):
        se, s0, s1, s2 = g["raw"][t]
        d1, d2 = (se / N) / 2, (s0 * n0 / P) / 2
        want = np.array([d1 * 2 / n0, d2 * 2 / n0, (d1 + d2) * 2 / n0, s1 / P, s2 / P]) / 4
        assert relnorm(D[ti], want) < TOL
    Es = sim.field_history()  # Es[:,ti] = mean of E over the window
    assert Es.shape == (N, 2)
    assert relnorm(Es[:, 0], g["E"][:4].mean(axis=0)) < TOL and relnorm(Es[:, 1], g["E"][4:8].mean(axis=0)) < TOL
    # stepping in one call gives the same state
    s2_ = pg.ngp_1d2v(T=64, TO=16)
    s2_.set_particles(g["x0"], g["vx0"], vy=g["vy0"])
    s2_.step(8)
    x2, vx2, vy2 = s2_.particles()
    assert np.array_equal(x2, x) and np.array_equal(vx2, vx) and np.array_equal(vy2, vy)
    # Boris rotation alone conserves |v| (E = 0 field: uniform plasma of identical particles has rho = const)
    a, b = oracle.boris_1d2v(0.3, -0.2, 0.0, 0.7, 0.05)
    assert abs(a * a + b * b - 0.13) < 1e-16


def test_ngp1d2v2s_steps(pg, oracle):
    """Two-species magnetised code (src/NGP1D2V2S.jl): golden fixture, diagnostics with the mass-weighted sums, and a
    live-oracle run with another mass ratio."""
    g = golden("ngp1d2v2s")
    N, P, M = int(g["N"]), int(g["P"]), float(g["M"])
    sim = pg.ngp_1d2v_2s(T=64, TO=16)  # window T/TO = 4 steps
    assert (sim.cfg.N, sim.cfg.P, sim.cfg.diag_every, sim.cfg.half_width, sim.count) == (256, 2048, 4, 7, 4096)
    assert sim.cfg.dt == float(g["dt"]) and sim.cfg.B0 == float(g["B0"]) and sim.cfg.w == float(g["w"]) and sim.cfg.mass_ratio == M
    sim.set_particles(g["x0"], g["vx0"], vy=g["vy0"])
    for t in range(8):
        sim.step(1)
        rho, E = sim.fields()
        # the two species neutralise each other (exactly so at the first step: both are loaded on a uniform lattice and the
        # integer charge grid sums to exactly 0, the oracle to 1e-17): compare on the scale of one species' density
        assert np.abs(rho - g["rho"][t]).max() < TOL * float(g["n0"])
        assert np.abs(E - g["E"][t]).max() < max(1e-10 * np.abs(g["E"][t]).max(), 1e-13 * float(g["n0"]))
    x, vx, vy = sim.particles()
    assert relnorm(x, g["x"]) < TOL and relnorm(vx, g["vx"]) < TOL and relnorm(vy, g["vy"]) < TOL
    D, _ = sim.diagnostics()
    assert D.shape == (2, 5)
    n0 = float(g["n0"])
    for ti, t in enumerate((3, 7)):  # NGP1D2V2S.jl:50-53,56
        se, s0, s1, s2 = g["raw"][t]
        d1, d2 = (se / N) / 2, (s0 * n0 / P) / 2
        want = np.array([d1 * 2 / n0, d2 * 2 / n0, (d1 + d2) * 2 / n0, s1 / P, s2 / P]) / 4
        assert np.abs(D[ti] - want).max() < 1e-10 * np.abs(want).max()
    Es = sim.field_history()
    assert relnorm(Es[:, 0], g["E"][:4].mean(axis=0)) < TOL and relnorm(Es[:, 1], g["E"][4:8].mean(axis=0)) < TOL
    # another mass ratio and size against the live oracle
    rng = np.random.default_rng(4)
    N2, P2, M2 = 128, 3000, 100.0
    s2_ = pg.ngp_1d2v_2s(N=N2, P=P2, T=16, TO=16, M=M2)
    x0, vx0, vy0 = rng.random(2 * P2), s2_.vth * rng.standard_normal(2 * P2), s2_.vth * rng.standard_normal(2 * P2)
    s2_.set_particles(x0, vx0, vy=vy0)
    s2_.step(3)
    xo, vxo, vyo = x0.copy(), vx0.copy(), vy0.copy()
    for _ in range(3):
        ro, Eo, _ = oracle.step_1d2v2s(xo, vxo, vyo, N2, 7, s2_.cfg.dt, s2_.cfg.B0, s2_.cfg.w, M2)
    xs, vxs, vys = s2_.particles()
    rs, Es2 = s2_.fields()
    assert relnorm(xs, xo) < TOL and relnorm(vxs, vxo) < TOL and relnorm(vys, vyo) < TOL and relnorm(Es2, Eo) < TOL
    assert np.abs(rs - ro).max() < TOL * n0


@pytest.mark.parametrize("variant", ["stream", "stream11_512"])
def test_2d3v_kernel_variants_match(pg, oracle, monkeypatch, variant):
    """The tile-sorted 2D path has two particle kernels -- particles_2d3v_stream (slice streaming through per-warp cp.async
    rings with replicated shared-memory windows, the default; also built without the replicas) and particles_2d3v_tiled
    (8192-particle work items, plain loads, PICGOLF_2D_KERNEL=tiled):
    all must give the same physics (to round-off: the counting sort ranks with atomics, so the split of a tile into work
    items, and with it the rounding of the window sums, differs from run to run).  P is odd so that work items end in ragged rows."""
    NX = NY = 64
    P = (1 << 18) + 3
    rng = np.random.default_rng(51)
    res = []
    for kern in ("tiled", variant):
        monkeypatch.setenv("PICGOLF_2D_KERNEL", kern)
        sim = pg.electrostatic_2d3v(NX=NX, NY=NY, P=P, T=8, NS=1, deposit_mode=pg.DEPOSIT_SORTED, sort_every=3)
        if not res:
            st = [1 - rng.random(P), 1 - rng.random(P)] + [rng.standard_normal(P) * sim.vth / np.sqrt(2) for _ in range(3)]
        sim.set_particles(st[0], st[2], y=st[1], vy=st[3], vz=st[4])
        sim.step(7)
        res.append((sim.particles(), sim.fields(), sim.diagnostics()[0]))
    for a, b in zip(res[0][0], res[1][0]):
        assert relnorm(b, a) < TOL
    assert relnorm(res[1][1][0], res[0][1][0]) < TOL and relnorm(res[1][1][1], res[0][1][1]) < TOL
    assert relnorm(res[1][2][:, :3], res[0][2][:, :3]) < TOL


# =============================================================================================
# picgolf_step_streamed: the pipelined set -> step -> get
# =============================================================================================
@pytest.mark.parametrize("scheme", ["fixedpoint", "fixedpoint_big", "ngp", "gauss_leapfrog", "2d3v"])
def test_step_streamed_matches_plain_calls(pg, scheme):
    """Seven different particle states pushed back to back through picgolf_step_streamed (three buffer sets, upload / step /
    download on three streams) give bit for bit what set_particles -> step(1) -> get_particles gives one state at a time,
    diagnostics rows included; afterwards the handle still serves the plain calls."""
    rng = np.random.default_rng(77)
    if scheme == "fixedpoint":
        mk = lambda: pg.gaussian_fixed_point(N=128, P=4096, T=16, W=400.0)
        P, ncomp = 4096, 2
    elif scheme == "fixedpoint_big":  # AUTO resolves to the sorted / polynomial path: a single step from a fresh state runs any-order
        mk = lambda: pg.gaussian_fixed_point(N=256, P=1 << 22, T=16, W=400.0)
        P, ncomp = 1 << 22, 2
    elif scheme == "ngp":
        mk = lambda: pg.ngp_fourier(N=4096, P=(1 << 21) + 5, NT=16, W=256.0)
        P, ncomp = (1 << 21) + 5, 2
    elif scheme == "gauss_leapfrog":
        mk = lambda: pg.gaussian(NX=128, NP=8192, NT=16)
        P, ncomp = 8192, 2
    else:
        mk = lambda: pg.electrostatic_2d3v(NX=64, NY=128, P=1 << 18, T=16, NS=1)
        P, ncomp = 1 << 18, 5
    a, b = mk(), mk()
    states = []
    for i in range(7):
        if ncomp == 2:
            states.append([rng.random(P), np.where(rng.random(P) < 0.5, -1.0, 1.0) + 0.01 * rng.standard_normal(P)])
        else:
            states.append([1 - rng.random(P), 1 - rng.random(P)] + [rng.standard_normal(P) * a.vth / math.sqrt(2) for _ in range(3)])
    outs = [[np.full(P, np.nan) for _ in range(ncomp)] for _ in states]
    for st, out in zip(states, outs):
        b.step_streamed(st, out)
    b.synchronize()
    for st, out in zip(states, outs):
        if ncomp == 2:
            a.set_particles(st[0], st[1])
        else:
            a.set_particles(st[0], st[2], y=st[1], vy=st[3], vz=st[4])
        a.step(1)
        for got, want in zip(out, a.particles()):
            assert np.array_equal(got, want)
    for fa, fb in zip(a.fields(), b.fields()):  # fields of the last state
        assert np.array_equal(fa, fb)
    Da, Db = a.diagnostics()[0], b.diagnostics()[0]
    assert Db.shape[0] == 7 and np.array_equal(Db[-1], Da[-1])  # one row per streamed call; the plain handle restarts its trace at every set
    # back to the plain calls on the streamed handle
    if ncomp == 2:
        b.set_particles(states[0][0], states[0][1])
    else:
        b.set_particles(states[0][0], states[0][2], y=states[0][1], vy=states[0][3], vz=states[0][4])
    b.step(1)
    for got, want in zip(b.particles(), outs[0]):
        assert np.array_equal(got, want)
