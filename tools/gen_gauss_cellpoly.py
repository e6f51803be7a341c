#!/usr/bin/env python3
"""Generate particleincellcodegolf.jl_b200/csrc/gauss_cellpoly.inc: the erf-shape stencil as one polynomial per EIGHTH of a cell.

For a power-of-two grid the 13 weights of the reference's stencil d(c) (src/GaussianFixedPoint.jl:4-5) depend only
on delta = c*N - k, k = round(c*N) the stencil centre (tools/gen_gauss_coeffs.py).  The cell is cut into NSUB = 8
intervals: with y = c*N*NSUB, interval m = round(y) holds the particles with u = y - m in [-1/2, 1/2]; it belongs to
centre k = (m + NSUB/2) >> 3 and sub-interval s = (m + NSUB/2) & 7, and delta = (s - NSUB/2 + u)/NSUB.  On each
sub-interval every weight is written in the same basis,

    W_j(delta) = erf(j + 1/2 - delta)/2 - erf(j - 1/2 - delta)/2 = sum_n CWS[s][j+6][n] * u^n,    |u| <= 1,

which makes both particle<->grid operations of a sweep linear in the powers of u (pg_kernels_poly.cuh):

    gather   sum_j E[k+j] * W_j   = sum_n u^n * G[m][n],        G[m][n] = sum_j CWS[s][j][n] * E[k+j]     (per interval, per solve)
    deposit  rho[i] = sum_p W_{i-k_p} = sum_j sum_s sum_n CWS[s][j][n] * M[m(i-j, s)][n],   M[m][n] = sum_{p in interval m} u_p^n

so a particle costs one degree-10 Horner evaluation (gather) and 10 power accumulations (deposit) instead of two
13-weight stencil evaluations -- and a third less than the one-polynomial-per-cell form of degree 16 this replaces.
The fit covers |u| <= 1, TWICE the interval: a lane of the particle pass keeps depositing into the interval it is
in until a particle lies a whole interval away from its centre (hysteresis), so the sub-cell bins of the sorted
particle order (1/32 cell wide) never make a lane alternate between two moment sets.  For |delta| > 1/2 the 13
weights are still those of the analytic shape around centre k; the cell the reference would add at the far end
weighs erfc(5.9)/2 < 3e-17 there, below the binary64 resolution of the weights.
Degree 10 reproduces every weight to <= 9e-17 (binary64 ulp(0.5) = 1.1e-16); sum_j CWS[s][j][n] = [n == 0] in the
60-digit fit (charge conservation) because the centre row is formed as the complement of the others.
Run: python tools/gen_gauss_cellpoly.py [--check]
"""
import os
import sys

import mpmath as mp

mp.mp.dps = 60
HALF = mp.mpf(1) / 2
DEG = 10
HW = 6
NSUB = 8
RANGE = mp.mpf(1)  # the fit covers |u| <= RANGE (PG_CWS_UMAX): twice the interval
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "particleincellcodegolf.jl_b200", "csrc", "gauss_cellpoly.inc")


def W(j, d):
    return (mp.erf(j + HALF - d) - mp.erf(j - HALF - d)) / 2


def delta(s, u):
    return (mp.mpf(s - NSUB // 2) + u) / NSUB


def cheb_monomial(f, n):
    """Degree-n Chebyshev interpolant of f on [-RANGE, RANGE] as monomial coefficients."""
    nodes = [RANGE * mp.cos(mp.pi * (2 * k + 1) / (2 * (n + 1))) for k in range(n + 1)]
    V = mp.matrix(n + 1, n + 1)
    y = mp.matrix(n + 1, 1)
    for i, x in enumerate(nodes):
        for k in range(n + 1):
            V[i, k] = x ** k
        y[i] = f(x)
    c = mp.lu_solve(V, y)
    return [c[i] for i in range(n + 1)]


def table():
    """CWS[s][j+HW][n] as mpf."""
    tab = []
    for s in range(NSUB):
        rows = {j: cheb_monomial(lambda u, j=j, s=s: W(j, delta(s, u)), DEG) for j in range(-HW, HW + 1) if j != 0}
        # centre row = complement: sum_j W_j = 1 (the +-7 edges are exactly +-1/2 in binary64, pg_gauss.cuh)
        rows[0] = [(1 if n == 0 else 0) - sum(rows[j][n] for j in rows) for n in range(DEG + 1)]
        tab.append([[rows[j][n] for n in range(DEG + 1)] for j in range(-HW, HW + 1)])
    return tab


def as_floats(tab):
    return [[[float(c) for c in row] for row in sub] for sub in tab]


def check(tabf):
    """Max |sum_n CWS u^n - W_j| over |u| <= 1 with the coefficients rounded to binary64 (the centre row against the
    complement of the other weights, which is what 'the stencil sums to one' makes it), and the charge defect."""
    worst = mp.mpf(0)
    for s, sub in enumerate(tabf):
        for i in range(129):
            u = (mp.mpf(2 * i) / 128 - 1) * RANGE
            d = delta(s, u)
            for jj, row in enumerate(sub):
                j = jj - HW
                exact = W(j, d) if j != 0 else 1 - sum(W(q, d) for q in range(-HW, HW + 1) if q != 0)
                approx = sum(mp.mpf(c) * u ** n for n, c in enumerate(row))
                worst = max(worst, abs(approx - exact))
    defect = max(abs(sum(mp.mpf(sub[jj][n]) for jj in range(2 * HW + 1)) - (1 if n == 0 else 0)) for sub in tabf for n in range(DEG + 1))
    return worst, defect


def main():
    global NSUB, DEG, OUT, RANGE
    # other tables for A/B builds: --nsub 16 --deg 8 --range 0.875 --out <path> (-DPG_CWS_ALT: 2.7 % faster on cold beams at 2^28
    # particles -- 41 instead of 47 FP64 instructions per particle-sweep -- but 10 / 28 % slower at thermal spreads of 0.3 / 1.0 of the
    # beam speed, where the narrower intervals turn more deposits into outliers; measured on B200, profiles/r2_s_*)
    for flag, conv in (("--nsub", int), ("--deg", int), ("--range", mp.mpf), ("--out", str)):
        if flag in sys.argv:
            val = conv(sys.argv[sys.argv.index(flag) + 1])
            if flag == "--nsub": NSUB = val
            elif flag == "--deg": DEG = val
            elif flag == "--range": RANGE = val
            else: OUT = val
    tabf = as_floats(table())
    worst, defect = check(tabf)
    assert worst < mp.mpf("1.1e-16") or "--out" in sys.argv, worst
    out = []
    out.append("// GENERATED by tools/gen_gauss_cellpoly.py -- do not edit.")
    out.append("// Sub-cell polynomial form of the erf-shape stencil: W_j(delta) = sum_n PG_CWS[s][j+6][n] * u^n on sub-interval s of a cell,")
    out.append(f"// delta = (s - {NSUB // 2} + u)/{NSUB}, fitted for |u| <= {mp.nstr(RANGE, 4)} (the interval is |u| <= 1/2), j = -6..6 (offset of the cell from the centre k), n = 0..{DEG}.")
    out.append(f"// max |fit - exact| over all weights (binary64 coefficients): {mp.nstr(worst, 3)}")
    out.append(f"// max |sum_j PG_CWS[s][j][n] - [n==0]|: {mp.nstr(defect, 3)}")
    out.append(f"#define PG_CWS_NSUB {NSUB}")
    out.append(f"#define PG_CWS_NC {DEG + 1}")
    out.append(f"#define PG_CWS_UMAX {mp.nstr(RANGE, 17)}")
    out.append(f"__constant__ double PG_CWS[{NSUB}][{2 * HW + 1}][{DEG + 1}] = {{")
    for sub in tabf:
        out.append("  {")
        for row in sub:
            out.append("    {" + ", ".join(repr(c) for c in row) + "},")
        out.append("  },")
    out.append("};")
    out.append("")
    text = "\n".join(out)
    if "--check" in sys.argv:
        ok = os.path.exists(OUT) and open(OUT).read() == text
        print("up to date" if ok else "STALE", "max err", mp.nstr(worst, 3), "charge defect", mp.nstr(defect, 3))
        sys.exit(0 if ok else 1)
    with open(OUT, "w") as fh:
        fh.write(text)
    print("\n".join(out[:6]))
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
