"""ctypes loader for the CPU parity oracle (oracle/picgolf_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, bench.py's cpu_baseline / --impl reference
legs and __graft_entry__.smoke().  The product package never imports this module.
Every wrapper mirrors one `oracle_*` C function; see the C file for the reference file:line.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpicgolf_oracle.so")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64 = C.c_int64
_d = C.c_double
_vp = C.c_void_p


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "picgolf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libpicgolf_oracle.so"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_jl_mod1.restype = _d
        L.oracle_jl_mod1.argtypes = [_d]
        L.oracle_ngp_index.restype = _i64
        L.oracle_ngp_index.argtypes = [_d, _i64]
        L.oracle_ngp_index_array.argtypes = [_dp, _i64, _i64, _ip]
        L.oracle_fft.argtypes = [_dp, _dp, _i64, C.c_int]
        L.oracle_dft_naive.argtypes = [_dp, _dp, _i64, C.c_int]
        L.oracle_solve1d.argtypes = [_dp, _i64, _dp]
        L.oracle_ngp_half_drift.argtypes = [_dp, _dp, _i64, _d]
        L.oracle_ngp_deposit.argtypes = [_dp, _i64, _i64, _d, _dp]
        L.oracle_ngp_kick.argtypes = [_dp, _dp, _dp, _i64, _i64, _d]
        L.oracle_ngp_step.argtypes = [_dp, _dp, _i64, _i64, _d, _d, _dp, _dp, _vp]
        L.oracle_gauss_stencil.argtypes = [_d, _i64, C.c_int, _ip, _dp]
        L.oracle_gauss_deposit.argtypes = [_dp, _dp, _i64, _i64, C.c_int, _d, _dp]
        L.oracle_gauss_gather.restype = _d
        L.oracle_gauss_gather.argtypes = [_dp, _d, _i64, C.c_int]
        L.oracle_isapprox.restype = C.c_int
        L.oracle_isapprox.argtypes = [_dp, _dp, _i64, _d, _d]
        L.oracle_fixedpoint_step.restype = C.c_int
        L.oracle_fixedpoint_step.argtypes = [_dp] * 7 + [_i64, _i64, C.c_int, _d, _d, _d, _d, _d, C.c_int, _vp, _vp]
        L.oracle_fixedpoint_run.argtypes = [_dp, _dp, _dp, _i64, _i64, C.c_int, _d, _d, _d, _d, _d, C.c_int, _i64, _vp, _vp]
        L.oracle_simpson_step.restype = C.c_int
        L.oracle_simpson_step.argtypes = [_dp] * 7 + [_i64, _i64, C.c_int, _d, _d, _d, _d, _d, C.c_int, _vp, C.c_int]
        L.oracle_area_stencil.argtypes = [_d, _i64, _ip, _dp]
        L.oracle_simpson_run.argtypes = [_dp, _dp, _dp, _i64, _i64, C.c_int, _d, _d, _d, _d, _d, C.c_int, _i64, _vp, _vp, C.c_int]
        L.oracle_boris_1d2v.argtypes = [_dp, _dp, _d, _d, _d]
        L.oracle_1d2v_step.argtypes = [_dp, _dp, _dp, _i64, _i64, C.c_int, _d, _d, _d, _dp, _dp, _vp]
        L.oracle_boris_1d2v_qm.argtypes = [_dp, _dp, _d, _d, _d, _d]
        L.oracle_1d2v2s_step.argtypes = [_dp, _dp, _dp, _i64, _i64, C.c_int, _d, _d, _d, _d, _dp, _dp, _vp]
        L.oracle_gauss_leapfrog_step.argtypes = [_dp, _dp, _i64, _i64, C.c_int, _d, _d, _dp, _dp, _vp]
        L.oracle_quiet_start.argtypes = [_i64, _i64, _i64, _dp, _dp]
        L.oracle_growth_slope.restype = _d
        L.oracle_growth_slope.argtypes = [_d]
        L.oracle_cic_g.argtypes = [_d, _i64, _ip, _dp]
        L.oracle_boris.argtypes = [_dp, _dp, _dp, _d, _d, _d, _d]
        L.oracle_cic_deposit.argtypes = [_dp, _dp, _i64, _i64, _i64, _d, _dp]
        L.oracle_cic_gather.argtypes = [_dp, _dp, _i64, _i64, _dp, _dp, _i64, _dp, _dp]
        L.oracle_solve2d.argtypes = [_dp, _i64, _i64, _dp, _dp]
        L.oracle_2d3v_step.argtypes = [_dp] * 5 + [_i64, _i64, _i64, _d, _d, _d, _dp, _dp, _dp, C.c_int]
        L.oracle_2d3v_diagnostics.argtypes = [_dp, _dp, _i64, _i64, _dp, _dp, _i64, _d, _dp]
        L.oracle_max_threads.restype = C.c_int
        _lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
        L.oracle_es_halton.restype = _d
        L.oracle_es_halton.argtypes = [_i64, _i64, _d]
        L.oracle_es_shape.restype = C.c_int
        L.oracle_es_shape.argtypes = [C.c_int, _d, _d, _lp, _dp]
        L.oracle_es_boris.argtypes = [_dp, _d, _d, _dp, _d, _d]
        L.oracle_es_loop.argtypes = [C.c_int, _lp, _ip, _dp, _dp, _dp] + [_dp] * 5 + [_i64, _i64, _d, _d, _d, _dp, _dp, C.c_int,
                                     _dp, _dp, _dp, _dp, C.c_int]
        L.oracle_es_diagnose.argtypes = [C.c_int, _lp, _dp, _dp, _dp, _dp, _dp, _i64, _i64, _dp, _dp]
        L.oracle_es_species.restype = _d
        L.oracle_es_species.argtypes = [_i64, _d, _d, _d, _d, _dp] + [_dp] * 5
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


# ---------------------------------------------------------------- Base semantics
def jl_mod1(x: float) -> float:
    return lib().oracle_jl_mod1(float(x))


def ngp_index(x, N: int) -> np.ndarray:
    """1-based NGP cell f(x) (NGPFourier.jl:3)."""
    x = _f64(np.atleast_1d(x))
    out = np.empty(x.size, dtype=np.int32)
    lib().oracle_ngp_index_array(x, x.size, N, out)
    return out


def fft(re, im, sign=-1):
    re, im = _f64(re).copy(), _f64(im).copy()
    lib().oracle_fft(re, im, re.size, sign)
    return re, im


def dft_naive(re, im, sign=-1):
    re, im = _f64(re).copy(), _f64(im).copy()
    lib().oracle_dft_naive(re, im, re.size, sign)
    return re, im


def solve1d(rho) -> np.ndarray:
    rho = _f64(rho)
    E = np.empty_like(rho)
    lib().oracle_solve1d(rho, rho.size, E)
    return E


# ---------------------------------------------------------------- NGP 1D1V
def ngp_half_drift(x, v, dt):
    lib().oracle_ngp_half_drift(x, v, x.size, dt)


def ngp_deposit(x, N, w) -> np.ndarray:
    n = np.empty(N)
    lib().oracle_ngp_deposit(_f64(x), x.size, N, w, n)
    return n


def ngp_kick(x, v, E, dt):
    lib().oracle_ngp_kick(x, v, _f64(E), x.size, E.size, dt)


def ngp_step(x, v, N, dt, w):
    """In place on x, v.  Returns (rho, E, raw[sumE2, sumv2, sumv])."""
    rho, E, raw = np.empty(N), np.empty(N), np.empty(3)
    lib().oracle_ngp_step(x, v, x.size, N, dt, w, rho, E, _ptr(raw))
    return rho, E, raw


# ---------------------------------------------------------------- Gaussian shape
def gauss_stencil(c: float, N: int, hw: int):
    idx = np.empty(2 * hw + 1, dtype=np.int32)
    wt = np.empty(2 * hw + 1)
    lib().oracle_gauss_stencil(float(c), N, hw, idx, wt)
    return idx, wt


def gauss_deposit(x, y, N, hw, scale) -> np.ndarray:
    r = np.empty(N)
    lib().oracle_gauss_deposit(_f64(x), _f64(y), x.size, N, hw, scale, r)
    return r


def gauss_gather(E, c, N, hw) -> np.ndarray:
    E = _f64(E)
    c = np.atleast_1d(_f64(c))
    L = lib()
    return np.array([L.oracle_gauss_gather(E, float(ci), N, hw) for ci in c])


def isapprox(F, E, rtol, atol=0.0) -> bool:
    return bool(lib().oracle_isapprox(_f64(F), _f64(E), F.size, rtol, atol))


class FixedPoint:
    """State holder mirroring GaussianFixedPoint.jl:1-6 / GaussianFixedPointQuiet.jl:1-7."""

    def __init__(self, x, v, N, dt, W, w=None, hw=6, rtol=1e-8, atol=0.0, max_sweeps=10):
        self.x, self.v = _f64(x).copy(), _f64(v).copy()
        self.P, self.N, self.dt, self.W, self.hw = self.x.size, N, dt, W, hw
        self.w = W / self.P * N if w is None else w
        self.rtol, self.atol, self.max_sweeps = rtol, atol, max_sweeps
        self.E, self.F, self.r = np.zeros(N), np.zeros(N), np.zeros(N)
        self.X, self.V = self.x.copy(), self.v.copy()

    def step(self):
        """One time step. Returns (D[t,1:4], raw sums, sweeps)."""
        D4, raw = np.empty(4), np.empty(3)
        s = lib().oracle_fixedpoint_step(self.x, self.v, self.E, self.X, self.V, self.F, self.r, self.P, self.N,
                                         self.hw, self.dt, self.W, self.w, self.rtol, self.atol, self.max_sweeps,
                                         _ptr(D4), _ptr(raw))
        return D4, raw, s

    def run(self, T):
        """T steps. Returns (D as T x 4 Fortran-ordered array, sweeps[T])."""
        D = np.zeros((T, 4), order="F")
        sw = np.zeros(T, dtype=np.int32)
        lib().oracle_fixedpoint_run(self.x, self.v, self.E, self.P, self.N, self.hw, self.dt, self.W, self.w,
                                    self.rtol, self.atol, self.max_sweeps, T, _ptr(D), _ptr(sw))
        return D, sw


class Simpson13(FixedPoint):
    """GaussianFixedPointQuietSimpson13.jl: E is 3 x N (rows E1 | E2 | E3 stored consecutively)."""

    def __init__(self, x, v, N, dt, W, w=None, hw=7, rtol=4 * np.finfo(float).eps, atol=0.0, max_sweeps=10, shape=0):
        """shape 0: Gaussian (GaussianFixedPointQuietSimpson13.jl); shape 1: area d(y) (AreaFixedPointQuietSimpson13.jl, l=1e-14)."""
        super().__init__(x, v, N, dt, W, w=w, hw=hw, rtol=rtol, atol=atol, max_sweeps=max_sweeps)
        self.E, self.F = np.zeros(3 * N), np.zeros(3 * N)
        self.shape = shape

    def step(self):
        D4 = np.empty(4)
        s = lib().oracle_simpson_step(self.x, self.v, self.E, self.X, self.V, self.F, self.r, self.P, self.N, self.hw,
                                      self.dt, self.W, self.w, self.rtol, self.atol, self.max_sweeps, _ptr(D4), self.shape)
        return D4, None, s

    def run(self, T):
        D = np.zeros((T, 4), order="F")
        sw = np.zeros(T, dtype=np.int32)
        lib().oracle_simpson_run(self.x, self.v, self.E, self.P, self.N, self.hw, self.dt, self.W, self.w, self.rtol,
                                 self.atol, self.max_sweeps, T, _ptr(D), _ptr(sw), self.shape)
        return D, sw


def area_stencil(y, N):
    idx, wt = np.empty(2, dtype=np.int32), np.empty(2)
    lib().oracle_area_stencil(float(y), N, idx, wt)
    return idx, wt


def gauss_leapfrog_step(x, v, N, hw, dt, scale):
    rho, E, raw = np.empty(N), np.empty(N), np.empty(3)
    lib().oracle_gauss_leapfrog_step(x, v, x.size, N, hw, dt, scale, rho, E, _ptr(raw))
    return rho, E, raw


def boris_1d2v(vx, vy, E, B, dt):
    a, b = np.array([vx], dtype=np.float64), np.array([vy], dtype=np.float64)
    lib().oracle_boris_1d2v(a, b, E, B, dt)
    return a[0], b[0]


def step_1d2v(x, vx, vy, N, hw, dt, B0, w):
    """src/NGP1D2V.jl:40-45,55 in place on x, vx, vy.  Returns (rho, E, raw[sumE2, sum v^2, sum vx, sum vy])."""
    rho, E, raw = np.empty(N), np.empty(N), np.empty(4)
    lib().oracle_1d2v_step(x, vx, vy, x.size, N, hw, dt, B0, w, rho, E, _ptr(raw))
    return rho, E, raw


def boris_1d2v_qm(vx, vy, E, B, dt, q_m):
    """src/NGP1D2V2S.jl:5-11"""
    a, b = np.array([vx], dtype=np.float64), np.array([vy], dtype=np.float64)
    lib().oracle_boris_1d2v_qm(a, b, E, B, dt, q_m)
    return a[0], b[0]


def step_1d2v2s(x, vx, vy, N, hw, dt, B0, w, M):
    """src/NGP1D2V2S.jl:31-49 in place on the concatenated [species 1 | species 2] arrays (length 2P).
    Returns (rho, E, raw[sumE2, mass-weighted sum v^2, sum vx, sum vy])."""
    rho, E, raw = np.empty(N), np.empty(N), np.empty(4)
    lib().oracle_1d2v2s_step(x, vx, vy, x.size // 2, N, hw, dt, B0, w, M, rho, E, _ptr(raw))
    return rho, E, raw


def quiet_start(P, first=0, count=None):
    count = P - first if count is None else count
    x, v = np.empty(count), np.empty(count)
    lib().oracle_quiet_start(P, first, count, x, v)
    return x, v


def growth_slope(W) -> float:
    return lib().oracle_growth_slope(W)


# ---------------------------------------------------------------- 2D3V
def cic_g(z, NZ):
    idx, wt = np.empty(2, dtype=np.int32), np.empty(2)
    lib().oracle_cic_g(float(z), NZ, idx, wt)
    return idx, wt


def boris(vx, vy, vz, Ex, Ey, dt, B0):
    a, b, c = np.array([vx], dtype=np.float64), np.array([vy], dtype=np.float64), np.array([vz], dtype=np.float64)
    lib().oracle_boris(a, b, c, Ex, Ey, dt, B0)
    return a[0], b[0], c[0]


def cic_deposit(x, y, NX, NY, w):
    rho = np.empty(NX * NY)
    lib().oracle_cic_deposit(_f64(x), _f64(y), x.size, NX, NY, w, rho)
    return rho


def cic_gather(Ex, Ey, NX, NY, x, y):
    ex, ey = np.empty(x.size), np.empty(x.size)
    lib().oracle_cic_gather(_f64(Ex), _f64(Ey), NX, NY, _f64(x), _f64(y), x.size, ex, ey)
    return ex, ey


def solve2d(rho, NX, NY):
    Ex, Ey = np.empty(NX * NY), np.empty(NX * NY)
    lib().oracle_solve2d(_f64(rho), NX, NY, Ex, Ey)
    return Ex, Ey


def step_2d3v(x, y, vx, vy, vz, NX, NY, dt, B0, w, Ex, Ey, nthreads=1):
    """In place on particles and Ex, Ey (flat column-major NX*NY).  Returns rho."""
    rho = np.empty(NX * NY)
    lib().oracle_2d3v_step(x, y, vx, vy, vz, x.size, NX, NY, dt, B0, w, Ex, Ey, rho, nthreads)
    return rho


def diagnostics_2d3v(Ex, Ey, NX, NY, vx, vy, w):
    K = np.empty(5)
    lib().oracle_2d3v_diagnostics(_f64(Ex), _f64(Ey), NX, NY, _f64(vx), _f64(vy), vx.size, w, K)
    return K


def max_threads() -> int:
    return lib().oracle_max_threads()


# ---------------------------------------------------------------- PIC2D3V.jl ElectrostaticField path
SHAPE_NGP, SHAPE_AREA, SHAPE_BSPLINE0 = 0, 1, 10  # BSplineWeighting{N} = 10 + N


def es_halton(i, base, seed=0.0):
    return lib().oracle_es_halton(int(i), int(base), float(seed))


def es_shape(shape, z, NZ_Lz):
    """depositindicesfractions: (first index (1-based, unwrapped), fractions)."""
    j0, wt = np.zeros(1, dtype=np.int64), np.zeros(6)
    n = lib().oracle_es_shape(int(shape), float(z), float(NZ_Lz), j0, wt)
    return int(j0[0]), wt[:n].copy()


def es_boris(v, Ex, Ey, B, dt, q_m):
    v = np.array(v, dtype=np.float64)
    lib().oracle_es_boris(v, float(Ex), float(Ey), _f64(B), float(dt), float(q_m))
    return v


def es_species(P, vth, density, Lx=1.0, Ly=1.0):
    """Species(P, vth, density, shape; Lx, Ly): returns x, y, vx, vy, vz, weight (erfinv from scipy)."""
    from scipy.special import erfinv
    seed = 1 / np.sqrt(2.0)
    einv = np.concatenate([erfinv(2 * np.array([es_halton(i, b, seed) for i in range(P)]) - 1) for b in (5, 7, 9)])
    out = [np.empty(P) for _ in range(5)]
    w = lib().oracle_es_species(P, float(vth), float(density), float(Lx), float(Ly), _f64(einv), *out)
    return (*out, w)


class ESField:
    """State of PIC2D3V.ElectrostaticField + plasma + ElectrostaticDiagnostics, stepped by loop!/diagnose!
    (PIC2D3V.jl:530-581,1301-1330).  species: list of dicts(x,y,vx,vy,vz,charge,mass,weight,shape)."""

    def __init__(self, species, NX, NY, Lx, Ly, dt, B, NT, ntskip=1, ngskip=1, accumulate=True, nthreads=1):
        self.NX, self.NY, self.Lx, self.Ly, self.dt = NX, NY, float(Lx), float(Ly), float(dt)
        self.B = _f64(B)
        self.sP = np.array([len(s["x"]) for s in species], dtype=np.int64)
        self.sshape = np.array([s["shape"] for s in species], dtype=np.int32)
        self.scharge = _f64([s["charge"] for s in species])
        self.smass = _f64([s["mass"] for s in species])
        self.sweight = _f64([s["weight"] for s in species])
        cat = lambda k: _f64(np.concatenate([np.asarray(s[k], dtype=np.float64) for s in species])).copy()
        self.x, self.y, self.vx, self.vy, self.vz = cat("x"), cat("y"), cat("vx"), cat("vy"), cat("vz")
        self.hn = (NX + 6) * (NY + 6)
        self.Exy = np.zeros(2 * self.hn)
        self.rho, self.Ex, self.Ey, self.phir = (np.zeros(NX * NY) for _ in range(4))
        self.accumulate, self.nthreads = int(bool(accumulate)), nthreads
        self.ntskip, self.ngskip, self.t, self.ti = ntskip, ngskip, 0, 0
        ND = NT // ntskip
        self.scalars = np.zeros((ND, 8))  # kinetic, field, pmom[3], cmom[3]
        nxd, nyd = NX // ngskip, NY // ngskip
        self.Exs, self.Eys, self.phis = (np.zeros((nxd, nyd, ND), order="F") for _ in range(3))

    def exy_interior(self):
        """Exy[c, 1:NX, 1:NY] as two flat column-major NX*NY arrays."""
        NX, NY = self.NX, self.NY
        g = self.Exy.reshape(2, NY + 6, NX + 6)
        return g[0, 3:NY + 3, 3:NX + 3].ravel().copy(), g[1, 3:NY + 3, 3:NX + 3].ravel().copy()

    def step(self):
        """loop!(plasma, field, to, t, _); diagnose!(diagnostics, field, plasma, t, to)   2D3V.jl:123-126"""
        L = lib()
        L.oracle_es_loop(len(self.sP), self.sP, self.sshape, self.scharge, self.smass, self.sweight, self.x, self.y, self.vx,
                         self.vy, self.vz, self.NX, self.NY, self.Lx, self.Ly, self.dt, self.B, self.Exy, self.accumulate,
                         self.rho, self.Ex, self.Ey, self.phir, self.nthreads)
        t = self.t
        if t % self.ntskip == 0:
            self.ti += 1
            if self.ti <= self.scalars.shape[0]:
                L.oracle_es_diagnose(len(self.sP), self.sP, self.smass, self.sweight, self.vx, self.vy, self.vz, self.NX, self.NY,
                                     self.Exy, self.scalars[self.ti - 1])
        if 1 <= self.ti <= self.Exs.shape[2]:
            g = self.ngskip
            sub = lambda a: a.reshape(self.NY, self.NX).T[::g, ::g]
            self.Exs[:, :, self.ti - 1] += sub(self.Ex) / self.ntskip
            self.Eys[:, :, self.ti - 1] += sub(self.Ey) / self.ntskip
            self.phis[:, :, self.ti - 1] += sub(self.phir) / self.ntskip
        self.t += 1
