"""CPU checks of the sub-cell polynomial form of the erf stencil (tools/gen_gauss_cellpoly.py): the generated table is
current, reproduces the reference's weights ff(i,c) (src/GaussianFixedPoint.jl:4-5) on every sub-interval including the
hysteresis range |u| <= 1, and the moment form of the deposit / polynomial form of the gather (pg_kernels_poly.cuh:
interval m = round(c*N*8), centre k = (m+4)>>3, sub-interval s = (m+4)&7) are the same linear maps as the per-particle stencil."""
import os
import re
import subprocess
import sys

import numpy as np
from scipy.special import erf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "particleincellcodegolf.jl_b200", "csrc", "gauss_cellpoly.inc")
_text = open(INC).read()
NSUB = int(re.search(r"#define PG_CWS_NSUB (\d+)", _text).group(1))   # 8 intervals per cell
NC = int(re.search(r"#define PG_CWS_NC (\d+)", _text).group(1))       # degree 10
UMAX = float(re.search(r"#define PG_CWS_UMAX ([0-9.eE+-]+)", _text).group(1))  # fitted on |u| <= 1
SUBLG = NSUB.bit_length() - 1


def load_table():
    text = _text
    body = text[text.index(f"PG_CWS[{NSUB}][13][{NC}]"):]
    rows = re.findall(r"\{([^{}]+)\},", body)
    tab = np.array([[float(v) for v in r.split(",")] for r in rows]).reshape(NSUB, 13, NC)
    return tab


def test_table_is_current():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_gauss_cellpoly.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def intervals(c, N):
    """Interval m, centre k (Julia index), sub-interval s and u of stencil centres c -- the kernel's arithmetic."""
    y = c * N * NSUB  # exact for power-of-two N
    m = np.rint(y).astype(np.int64)
    return m, (m + NSUB // 2) >> SUBLG, (m + NSUB // 2) & (NSUB - 1), y - m


def test_weights_match_reference_expression():
    CWS = load_table()
    N = 4096
    rng = np.random.default_rng(0)
    c = rng.random(4000)
    m, k, s, u = intervals(c, N)
    assert np.abs(u).max() <= 0.5
    # the interval's centre k is the reference's round(c*N) except within half an interval above a cell edge
    # (c*N = k + 1/2 - 1/16 .. k + 1/2 belongs to sub-interval 0 of centre k+1: delta in [-9/16, -7/16])
    ctr = np.rint(c * N)
    assert set(np.unique(k - ctr)) <= {0, 1}
    delta = c * N - k
    assert np.allclose(delta, (s - NSUB // 2 + u) / NSUB, atol=1e-12)
    for jj in range(13):
        j = jj - 6
        got = np.array([np.polynomial.polynomial.polyval(ui, CWS[si, jj]) for ui, si in zip(u, s)])
        # exact-delta comparison (what the CUDA kernels evaluate): the analytic shape around centre k
        ref = (erf(j + 0.5 - delta) - erf(j - 0.5 - delta)) / 2
        assert np.abs(got - ref).max() < 4e-16
        # the reference's expression ff(i,c) for the same grid cell i = k + j (argument rounding N*eps*|g| ~ 4.5e-13 at N=4096)
        i = k + j
        ref2 = erf(((i + 0.5) / N - c) * N) / 2 - erf(((i - 0.5) / N - c) * N) / 2
        assert np.abs(got - ref2).max() < 2e-12
    import math
    for si in range(NSUB):  # charge conservation: sum_j CWS[s][j][n] = [n == 0] (exactly rounded sums of the binary64 table)
        for n in range(NC):
            assert abs(math.fsum(CWS[si, :, n]) - (1.0 if n == 0 else 0.0)) < 1.2e-16


def test_hysteresis_range_is_covered():
    """A warp keeps using interval m while |u| <= UMAX: the polynomials must hold there, and what the shifted stencil leaves out
    (the reference's 13th cell at the far end) must be below binary64 resolution."""
    CWS = load_table()
    u = np.linspace(-UMAX, UMAX, 401)
    worst = 0.0
    for s in range(NSUB):
        delta = (s - NSUB // 2 + u) / NSUB  # down to -1/2 - UMAX/NSUB
        for jj in range(13):
            j = jj - 6
            ref = (erf(j + 0.5 - delta) - erf(j - 0.5 - delta)) / 2
            if j == 0:
                ref = 1 - sum((erf(q + 0.5 - delta) - erf(q - 0.5 - delta)) / 2 for q in range(-6, 7) if q != 0)
            worst = max(worst, np.abs(np.polynomial.polynomial.polyval(u, CWS[s, jj]) - ref).max())
    assert worst < 4e-16
    # weight of the cell the reference's own 13-cell stencil would hold instead, at the far edge of the range
    assert (erf(6.5 + 0.5 + UMAX / NSUB) - erf(5.5 + 0.5 + UMAX / NSUB)) / 2 < 3e-17


def test_moment_deposit_and_poly_gather_are_the_stencil():
    CWS = load_table()
    N, P = 64, 5000
    rng = np.random.default_rng(1)
    c = rng.random(P)
    ctr = np.rint(c * N).astype(int)
    d = c * N - ctr
    E = rng.standard_normal(N)
    # direct stencil, as the reference writes it
    rho = np.zeros(N)
    g = np.zeros(P)
    for j in range(-6, 7):
        w = (erf(j + 0.5 - d) - erf(j - 0.5 - d)) / 2
        idx = (ctr + j - 1) % N  # Julia index i -> 0-based cell
        np.add.at(rho, idx, w)
        g += E[idx] * w
    # moment form; half of the particles deposit into a NEIGHBOURING interval (|u| up to 1), as a lane with hysteresis does
    m, _, _, u = intervals(c, N)
    shift = rng.integers(-1, 2, P) * (rng.random(P) < 0.5)
    shift = np.where(np.abs(u - shift) <= UMAX, shift, 0)
    m2, u2 = m + shift, u - shift
    M = np.zeros((N * NSUB, NC))
    row = m2 % (N * NSUB)
    for n in range(NC):
        np.add.at(M[:, n], row, u2 ** n)
    rho_m = np.zeros(N)
    G = np.zeros((N * NSUB, NC))
    for z in range(N):
        for s in range(NSUB):
            r = (NSUB * z + s + NSUB // 2) % (N * NSUB)  # cp_row_of
            for j in range(-6, 7):
                rho_m[(z + j) % N] += CWS[s, j + 6] @ M[r]
                G[r] += CWS[s, j + 6] * E[(z + j) % N]
    g_m = (G[m % (N * NSUB)] * u[:, None] ** np.arange(NC)).sum(axis=1)
    assert np.abs(rho_m - rho).max() < 1e-13 * np.abs(rho).max()
    assert np.abs(g_m - g).max() < 1e-14 * np.abs(E).max() * 13
    assert abs(rho_m.sum() - P) < 1e-9
