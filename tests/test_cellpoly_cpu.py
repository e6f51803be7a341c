"""CPU checks of the cell-polynomial form of the erf stencil (tools/gen_gauss_cellpoly.py): the generated table is
current, reproduces the reference's weights ff(i,c) (src/GaussianFixedPoint.jl:4-5), and the moment form of the
deposit / polynomial form of the gather (pg_kernels_poly.cuh) are the same linear maps as the per-particle stencil."""
import os
import re
import subprocess
import sys

import numpy as np
from scipy.special import erf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "particleincellcodegolf.jl_b200", "csrc", "gauss_cellpoly.inc")


def load_table():
    text = open(INC).read()
    body = text[text.index("PG_CW[13][17]"):]
    rows = re.findall(r"\{([^{}]+)\},", body)
    tab = np.array([[float(v) for v in r.split(",")] for r in rows])
    assert tab.shape == (13, 17)
    return tab


def test_table_is_current():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_gauss_cellpoly.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_weights_match_reference_expression():
    CW = load_table()
    N = 4096
    rng = np.random.default_rng(0)
    c = rng.random(2000)
    ctr = np.rint(c * N)
    t = 2 * (c * N - ctr)
    for jj in range(13):
        i = ctr + (jj - 6)
        ref = erf(((i + 0.5) / N - c) * N) / 2 - erf(((i - 0.5) / N - c) * N) / 2  # ff(i,c)
        got = np.polynomial.polynomial.polyval(t, CW[jj])
        # the reference's own argument rounding is N*eps*|g| ~ 4.5e-13 at N=4096; the table is exact in delta
        assert np.abs(got - ref).max() < 2e-12
    # exact-delta comparison (what the CUDA kernels evaluate): |error| <= 1e-16
    d = t / 2
    for jj in range(13):
        j = jj - 6
        ref = (erf(j + 0.5 - d) - erf(j - 0.5 - d)) / 2
        assert np.abs(np.polynomial.polynomial.polyval(t, CW[jj]) - ref).max() < 4e-16
    assert np.abs(CW.sum(axis=0) - np.eye(17)[0]).max() < 1e-16


def test_moment_deposit_and_poly_gather_are_the_stencil():
    CW = load_table()
    N, P = 64, 5000
    rng = np.random.default_rng(1)
    c = rng.random(P)
    ctr = np.rint(c * N).astype(int)
    d = c * N - ctr
    t = 2 * d
    E = rng.standard_normal(N)
    # direct stencil
    rho = np.zeros(N)
    g = np.zeros(P)
    for j in range(-6, 7):
        w = (erf(j + 0.5 - d) - erf(j - 0.5 - d)) / 2
        idx = (ctr + j - 1) % N  # Julia index i -> 0-based cell
        np.add.at(rho, idx, w)
        g += E[idx] * w
    # moment form
    M = np.zeros((N, 17))
    cell = (ctr - 1) % N
    for n in range(17):
        np.add.at(M[:, n], cell, t ** n)
    rho_m = np.zeros(N)
    G = np.zeros((N, 17))
    for j in range(-6, 7):
        rho_m += (np.roll(M, j, axis=0) * CW[j + 6]).sum(axis=1)  # cell i receives W_j from cell i-j
        G += np.outer(np.roll(E, -j), CW[j + 6])                   # G[c][n] = sum_j CW[j][n] E[c+j]
    g_m = (G[cell] * t[:, None] ** np.arange(17)).sum(axis=1)
    assert np.abs(rho_m - rho).max() < 1e-12 * np.abs(rho).max()
    assert np.abs(g_m - g).max() < 1e-14 * np.abs(E).max() * 13
