/*
 * picgolf_oracle.c -- CPU restatement of the per-timestep PIC loop of
 * jwscook/ParticleInCellCodeGolf.jl.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle and the timed CPU baseline ("port").  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  Nothing under particleincellcodegolf.jl_b200/ links, calls
 * or falls back to it.
 *
 * PARITY PINNING: the reference has no tests and no golden vectors and Julia is
 * not installed in the build container, so the reference cannot be executed here
 * ("parity unpinned" by reference tests).  The restatement is anchored by the
 * known answers the reference does hold: the analytic two-stream growth rate
 * overlaid in src/GaussianFixedPointQuiet.jl:19-20 / figs/GaussianFixedPointQuiet.jpg
 * and the conservation claims in README.md:46-47,76 (tests/test_oracle_known_answers.py).
 *
 * All citations are file:line relative to /root/reference/.  Expression order
 * follows the Julia source literally (Julia does not contract a*b+c into an FMA
 * unless @muladd is written, so compile this file with -ffp-contract=off).
 *
 * Third-party arithmetic the reference calls and that is not in its tree:
 *   FFTW.jl 1.7.1 / FFTW_jll 3.3.10 (Manifest.toml:315-325)  -> radix-2 FFT below
 *       (any correctly rounded-ish FFT agrees to ~log2(N)*eps norm-wise), plus an
 *       O(N^2) long-double DFT for non powers of two and as a cross-check.
 *   SpecialFunctions 2.2.0 -> OpenLibm_jll 0.8.1 erf (Manifest.toml:774-777,1054-1058)
 *       -> glibc erf (both < 1 ulp).
 *   Julia 1.8.5 Base: round (ties to even), float mod, mod1, bitreverse(::Int64),
 *       LinearAlgebra.isapprox(::Array, ::Array) (2-norm test).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define ORACLE_API __attribute__((visibility("default")))

static const double TWO_PI = 6.283185307179586476925286766559; /* Julia 2pi == 2*Float64(pi) */

/* ------------------------------------------------------------------ */
/* Julia Base semantics                                               */
/* ------------------------------------------------------------------ */

/* Julia mod(x::Float64, 1): base/float.jl `mod`: r = rem(x,y); r==0 -> copysign(r,y);
 * (r>0) xor (y>0) -> r+y; else r.   NGPFourier.jl:2, GaussianFixedPoint.jl:9. */
ORACLE_API double oracle_jl_mod1(double x)
{
    double r = fmod(x, 1.0);
    if (r == 0.0) return 0.0;
    if (r < 0.0) return r + 1.0; /* may round to exactly 1.0 for tiny negative x */
    return r;
}

/* Julia mod1(i, N) on integers: mod(i-1, N) + 1 with floored mod. GaussianFixedPoint.jl:5 */
static inline int64_t jl_imod1(int64_t i, int64_t N)
{
    int64_t m = (i - 1) % N;
    if (m < 0) m += N;
    return m + 1;
}

/* f(x) = Int(mod1(round(x*N), N))   NGPFourier.jl:3.  Returns the 1-based cell 1..N.
 * round = ties-to-even (rint under the default rounding mode); the mod1 is on Float64. */
ORACLE_API int64_t oracle_ngp_index(double x, int64_t N)
{
    double r = rint(x * (double)N);
    double m = fmod(r, (double)N);
    if (m != 0.0 && (m < 0.0)) m += (double)N; /* floored mod for negative r */
    if (m == 0.0) m = (double)N;               /* mod1: 0 -> N */
    return (int64_t)m;
}

ORACLE_API void oracle_ngp_index_array(const double *x, int64_t P, int64_t N, int32_t *idx1)
{
    for (int64_t j = 0; j < P; ++j) idx1[j] = (int32_t)oracle_ngp_index(x[j], N);
}

/* ------------------------------------------------------------------ */
/* FFT (stand-in for FFTW: unnormalised forward e^{-2 pi i nk/N}, inverse 1/N) */
/* ------------------------------------------------------------------ */

static int is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

static void dft_naive(double *re, double *im, int64_t N, int sign)
{
    long double *or_ = (long double *)malloc(sizeof(long double) * 2 * (size_t)N);
    long double *oi = or_ + N;
    for (int64_t k = 0; k < N; ++k) {
        long double sr = 0, si = 0;
        for (int64_t n = 0; n < N; ++n) {
            int64_t m = (k * n) % N;
            long double ang = (long double)sign * 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)N;
            long double c = cosl(ang), s = sinl(ang);
            sr += re[n] * c - im[n] * s;
            si += re[n] * s + im[n] * c;
        }
        or_[k] = sr; oi[k] = si;
    }
    for (int64_t k = 0; k < N; ++k) { re[k] = (double)or_[k]; im[k] = (double)oi[k]; }
    free(or_);
}

/* in-place iterative radix-2 DIT; sign=-1 forward, +1 backward (unnormalised) */
static void fft_pow2(double *re, double *im, int64_t N, int sign)
{
    int lg = 0; while (((int64_t)1 << lg) < N) ++lg;
    for (int64_t i = 0; i < N; ++i) {
        int64_t j = 0;
        for (int b = 0; b < lg; ++b) if (i & ((int64_t)1 << b)) j |= (int64_t)1 << (lg - 1 - b);
        if (j > i) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    double *wr = (double *)malloc(sizeof(double) * (size_t)N);
    double *wi = wr + N / 2;
    for (int64_t k = 0; k < N / 2; ++k) {
        long double ang = (long double)sign * 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)N;
        wr[k] = (double)cosl(ang); wi[k] = (double)sinl(ang);
    }
    for (int64_t len = 2; len <= N; len <<= 1) {
        int64_t half = len >> 1, step = N / len;
        for (int64_t s = 0; s < N; s += len)
            for (int64_t k = 0; k < half; ++k) {
                double c = wr[k * step], d = wi[k * step];
                double ar = re[s + k + half], ai = im[s + k + half];
                double tr = ar * c - ai * d, ti = ar * d + ai * c;
                re[s + k + half] = re[s + k] - tr; im[s + k + half] = im[s + k] - ti;
                re[s + k] += tr; im[s + k] += ti;
            }
    }
    free(wr);
}

static void fft1(double *re, double *im, int64_t N, int sign)
{
    if (N == 1) return;
    if (is_pow2(N)) fft_pow2(re, im, N, sign); else dft_naive(re, im, N, sign);
}

ORACLE_API void oracle_fft(double *re, double *im, int64_t N, int sign) { fft1(re, im, N, sign); }
ORACLE_API void oracle_dft_naive(double *re, double *im, int64_t N, int sign) { dft_naive(re, im, N, sign); }

/* ------------------------------------------------------------------ */
/* 1D spectral solve                                                  */
/* ------------------------------------------------------------------ */

/* E = real.(ifft((xi = fft(rho)./ik; xi[1]*=0; xi)))  with
 * ik = 2pi*im*vcat(1, 1:N/2, -N/2+1:-1)      NGPFourier.jl:3,5; GaussianFixedPoint.jl:3,8.
 * Slot 1 holds a dummy 1; the Nyquist slot N/2+1 holds +N/2.  z/(i b) = (Im z)/b - i (Re z)/b. */
ORACLE_API void oracle_solve1d(const double *rho, int64_t N, double *E)
{
    double *re = (double *)malloc(sizeof(double) * 2 * (size_t)N);
    double *im = re + N;
    for (int64_t n = 0; n < N; ++n) { re[n] = rho[n]; im[n] = 0.0; }
    fft1(re, im, N, -1);
    for (int64_t s = 0; s < N; ++s) {
        double kk = (s == 0) ? 1.0 : (s <= N / 2 ? (double)s : (double)(s - N));
        double b = TWO_PI * kk;
        double a_re = re[s], a_im = im[s];
        re[s] = a_im / b;
        im[s] = -a_re / b;
    }
    re[0] = 0.0; im[0] = 0.0; /* xi[1] *= 0 */
    fft1(re, im, N, +1);
    for (int64_t n = 0; n < N; ++n) E[n] = re[n] / (double)N;
    free(re);
}

/* ------------------------------------------------------------------ */
/* NGP 1D1V (NGPFourier.jl, NGPFourierWithDiagnostics.jl)              */
/* ------------------------------------------------------------------ */

/* u(): x .= mod.(x .+ v/2*dt, 1)        NGPFourier.jl:2  (v/2 first, then *dt) */
ORACLE_API void oracle_ngp_half_drift(double *x, const double *v, int64_t P, double dt)
{
    for (int64_t j = 0; j < P; ++j) x[j] = oracle_jl_mod1(x[j] + v[j] / 2 * dt);
}

/* n.*=0; for j in x; n[f(j)] += w; end   NGPFourier.jl:5 (sequential j = 1..P) */
ORACLE_API void oracle_ngp_deposit(const double *x, int64_t P, int64_t N, double w, double *n)
{
    for (int64_t i = 0; i < N; ++i) n[i] = 0.0;
    for (int64_t j = 0; j < P; ++j) n[oracle_ngp_index(x[j], N) - 1] += w;
}

/* v += E[f.(x)]*dt                       NGPFourier.jl:6 */
ORACLE_API void oracle_ngp_kick(const double *x, double *v, const double *E, int64_t P, int64_t N, double dt)
{
    for (int64_t j = 0; j < P; ++j) v[j] = v[j] + E[oracle_ngp_index(x[j], N) - 1] * dt;
}

/* One time step of NGPFourier.jl:5-6.  rho/E are outputs (E is the field of this step).
 * raw[0..2] = sum(E.^2), sum(v.^2), sum(v) after the kick (raw sums; see K below). */
ORACLE_API void oracle_ngp_step(double *x, double *v, int64_t P, int64_t N, double dt, double w,
                                double *rho, double *E, double *raw)
{
    oracle_ngp_half_drift(x, v, P, dt);
    oracle_ngp_deposit(x, P, N, w, rho);
    oracle_solve1d(rho, N, E);
    oracle_ngp_half_drift(x, v, P, dt);
    oracle_ngp_kick(x, v, E, P, N, dt);
    if (raw) {
        double se = 0, sv2 = 0, sv = 0;
        for (int64_t i = 0; i < N; ++i) se += E[i] * E[i];
        for (int64_t j = 0; j < P; ++j) { sv2 += v[j] * v[j]; sv += v[j]; }
        raw[0] = se; raw[1] = sv2; raw[2] = sv;
    }
}

/* ------------------------------------------------------------------ */
/* Gaussian erf shape (Gaussian.jl, GaussianFixedPoint.jl, ...Quiet.jl) */
/* ------------------------------------------------------------------ */

/* f(g,c) = erf((g-c)*N)/2                 GaussianFixedPoint.jl:4 */
static inline double shape_f(double g, double c, double N) { return erf((g - c) * N) / 2; }
/* ff(i,c) = f((i+0.5)/N,c) - f((i-0.5)/N,c)   GaussianFixedPoint.jl:4 */
static inline double shape_ff(int64_t i, double c, double N)
{
    return shape_f(((double)i + 0.5) / N, c, N) - shape_f(((double)i - 0.5) / N, c, N);
}

/* d(c) = ((mod1(i,N), ff(i,c)) for i in (-hw:hw) .+ Int(round(c*N)))   GaussianFixedPoint.jl:5
 * (hw = 6) and GaussianFixedPointQuiet.jl:6 (hw = 7).  idx1 is 1-based. */
ORACLE_API void oracle_gauss_stencil(double c, int64_t N, int hw, int32_t *idx1, double *wt)
{
    int64_t i0 = (int64_t)rint(c * (double)N);
    for (int k = 0; k <= 2 * hw; ++k) {
        int64_t i = i0 - hw + k;
        idx1[k] = (int32_t)jl_imod1(i, N);
        wt[k] = shape_ff(i, c, (double)N);
    }
}

/* rho(x,y) = (r.*=0; for j in d.((x.+y)./2); for k in j; r[k[1]] += k[2]*w; ...)
 * GaussianFixedPoint.jl:6 -- deposit at the midpoint (x+y)/2; scale = w (or w/dx for Gaussian.jl:7). */
ORACLE_API void oracle_gauss_deposit(const double *x, const double *y, int64_t P, int64_t N, int hw,
                                     double scale, double *r)
{
    int32_t idx[32]; double wt[32];
    for (int64_t i = 0; i < N; ++i) r[i] = 0.0;
    for (int64_t j = 0; j < P; ++j) {
        double c = (x[j] + y[j]) / 2;
        oracle_gauss_stencil(c, N, hw, idx, wt);
        for (int k = 0; k <= 2 * hw; ++k) r[idx[k] - 1] += wt[k] * scale;
    }
}

/* sum(k->E[k[1]]*k[2], d(c))  GaussianFixedPoint.jl:9 (left-to-right i = -hw..hw) */
ORACLE_API double oracle_gauss_gather(const double *E, double c, int64_t N, int hw)
{
    int32_t idx[32]; double wt[32];
    oracle_gauss_stencil(c, N, hw, idx, wt);
    double s = E[idx[0] - 1] * wt[0];
    for (int k = 1; k <= 2 * hw; ++k) s += E[idx[k] - 1] * wt[k];
    return s;
}

/* Julia LinearAlgebra.isapprox(F, E; rtol, atol) on arrays (generic.jl, Julia 1.8):
 * d = norm(F-E); isfinite(d) ? d <= max(atol, rtol*max(norm(F),norm(E))) : all(elementwise).
 * GaussianFixedPoint.jl:7, GaussianFixedPointQuiet.jl:8. */
ORACLE_API int oracle_isapprox(const double *F, const double *E, int64_t N, double rtol, double atol)
{
    double d2 = 0, f2 = 0, e2 = 0;
    for (int64_t i = 0; i < N; ++i) {
        double d = F[i] - E[i];
        d2 += d * d; f2 += F[i] * F[i]; e2 += E[i] * E[i];
    }
    double d = sqrt(d2);
    if (isfinite(d)) {
        double m = sqrt(f2) > sqrt(e2) ? sqrt(f2) : sqrt(e2);
        double tol = atol > rtol * m ? atol : rtol * m;
        return d <= tol;
    }
    for (int64_t i = 0; i < N; ++i) {
        double a = F[i], b = E[i];
        if (a == b) continue;
        if (!(isfinite(a) && isfinite(b))) return 0;
        double m = fabs(a) > fabs(b) ? fabs(a) : fabs(b);
        double tol = atol > rtol * m ? atol : rtol * m;
        if (!(fabs(a - b) <= tol)) return 0;
    }
    return 1;
}

/* One step of GaussianFixedPoint.jl:7-11 / GaussianFixedPointQuiet.jl:8-13.
 * State in: x,v (P), E (N; persists across steps).  Scratch: X,V (P), F,r (N).
 * Out: D4[0..3] = D[t,1:4]; raw[0..2] = sum(E.^2), sum(v.^2), sum(v).  Returns sweeps executed. */
ORACLE_API int oracle_fixedpoint_step(double *x, double *v, double *E, double *X, double *V, double *F,
                                      double *r, int64_t P, int64_t N, int hw, double dt, double W,
                                      double w, double rtol, double atol, int max_sweeps, double *D4,
                                      double *raw)
{
    int sweeps = 0;
    memcpy(X, x, sizeof(double) * (size_t)P);             /* X.=x */
    memcpy(V, v, sizeof(double) * (size_t)P);             /* V.=v */
    for (int64_t i = 0; i < N; ++i) F[i] = NAN;           /* F.*=NaN */
    for (int it = 0; it < max_sweeps; ++it) {             /* for _ in 0:9 */
        if (oracle_isapprox(F, E, N, rtol, atol)) break;  /* isapprox(F,E,rtol=l) && break */
        memcpy(F, E, sizeof(double) * (size_t)N);         /* F.=E */
        for (int64_t j = 0; j < P; ++j) x[j] = X[j] + (v[j] + V[j]) / 2 * dt; /* x.=X.+(v.+V)/2*dt */
        oracle_gauss_deposit(x, X, P, N, hw, w, r);       /* rho(x,X) */
        oracle_solve1d(r, N, E);
        for (int64_t j = 0; j < P; ++j)                    /* v[j]=V[j]+sum(...)*dt */
            v[j] = V[j] + oracle_gauss_gather(E, (x[j] + X[j]) / 2, N, hw) * dt;
        ++sweeps;
    }
    for (int64_t j = 0; j < P; ++j) x[j] = oracle_jl_mod1(x[j]); /* x.=mod.(x,1) */
    /* D[t,1:2].=(sum(E.^2)/N,sum(v.^2)*W/P)./2; D[t,3:4].=sum.((D[t,1:2],v/P)); D[t,1:3].*=2/W */
    double se = 0, sv2 = 0, sv = 0, svp = 0;
    for (int64_t i = 0; i < N; ++i) se += E[i] * E[i];
    for (int64_t j = 0; j < P; ++j) { sv2 += v[j] * v[j]; sv += v[j]; svp += v[j] / (double)P; }
    if (D4) {
        double d1 = (se / (double)N) / 2, d2 = (sv2 * W / (double)P) / 2;
        double d3 = d1 + d2;
        double s = 2 / W;
        D4[0] = d1 * s; D4[1] = d2 * s; D4[2] = d3 * s; D4[3] = svp;
    }
    if (raw) { raw[0] = se; raw[1] = sv2; raw[2] = sv; }
    return sweeps;
}

/* Run T steps; D is T x 4 column-major (Julia layout); sweeps[t] optional. */
ORACLE_API void oracle_fixedpoint_run(double *x, double *v, double *E, int64_t P, int64_t N, int hw,
                                      double dt, double W, double w, double rtol, double atol,
                                      int max_sweeps, int64_t T, double *D, int32_t *sweeps)
{
    double *X = (double *)malloc(sizeof(double) * (size_t)(2 * P + 2 * N));
    double *V = X + P, *F = V + P, *r = F + N;
    memcpy(F, E, sizeof(double) * (size_t)N);
    for (int64_t t = 0; t < T; ++t) {
        double d4[4];
        int s = oracle_fixedpoint_step(x, v, E, X, V, F, r, P, N, hw, dt, W, w, rtol, atol, max_sweeps, d4, NULL);
        if (D) for (int c = 0; c < 4; ++c) D[c * T + t] = d4[c];
        if (sweeps) sweeps[t] = s;
    }
    free(X);
}

/* d(y)=(i=Int(mod1(ceil(y*N),N));o=ceil(y*N)-y*N;((i,1-o),(mod1(i-1,N),o)))   AreaFixedPointQuietSimpson13.jl:5 */
ORACLE_API void oracle_area_stencil(double y, int64_t N, int32_t *idx1, double *wt)
{
    double ce = ceil(y * (double)N);
    double m = fmod(ce, (double)N);
    if (m < 0.0) m += (double)N;
    if (m == 0.0) m = (double)N;
    int64_t i = (int64_t)m;
    double o = ce - y * (double)N;
    idx1[0] = (int32_t)i; wt[0] = 1 - o;
    idx1[1] = (int32_t)jl_imod1(i - 1, N); wt[1] = o;
}

/* shape 0: Gaussian d(c) of half width hw; shape 1: area d(y).  deposit r[k[1]] += k[2]*w, gather sum(E[k[1]]*k[2]). */
static void shape_deposit(int shape, const double *x, const double *y, int64_t P, int64_t N, int hw, double w, double *r)
{
    if (shape == 0) { oracle_gauss_deposit(x, y, P, N, hw, w, r); return; }
    for (int64_t i = 0; i < N; ++i) r[i] = 0.0;
    for (int64_t j = 0; j < P; ++j) {
        int32_t idx[2]; double wt[2];
        oracle_area_stencil((x[j] + y[j]) / 2, N, idx, wt);
        r[idx[0] - 1] += wt[0] * w;
        r[idx[1] - 1] += wt[1] * w;
    }
}
static double shape_gather(int shape, const double *E, double c, int64_t N, int hw)
{
    if (shape == 0) return oracle_gauss_gather(E, c, N, hw);
    int32_t idx[2]; double wt[2];
    oracle_area_stencil(c, N, idx, wt);
    return E[idx[0] - 1] * wt[0] + E[idx[1] - 1] * wt[1];
}

/* One step of GaussianFixedPointQuietSimpson13.jl:8-18 (Simpson-1/3 time quadrature of E: three field
 * solves per sweep).  E is 3 x N stored row by row here (E1 | E2 | E3), F likewise; the convergence test is the
 * Frobenius-norm isapprox over the whole 3 x N matrix.  Returns sweeps; D4 as in the plain fixed point but with
 * sum(E[end,:].^2) (line 18).  shape 1 runs AreaFixedPointQuietSimpson13.jl:7-17 (same schedule, d(y) of its line 5). */
ORACLE_API int oracle_simpson_step(double *x, double *v, double *E3N, double *X, double *V, double *F3N, double *r,
                                   int64_t P, int64_t N, int hw, double dt, double W, double w, double rtol, double atol,
                                   int max_sweeps, double *D4, int shape)
{
    int sweeps = 0;
    double *E1 = E3N, *E2 = E3N + N, *E3 = E3N + 2 * N;
    memcpy(X, x, sizeof(double) * (size_t)P);
    memcpy(V, v, sizeof(double) * (size_t)P);
    for (int64_t i = 0; i < 3 * N; ++i) F3N[i] = NAN;                         /* F.*=NaN */
    shape_deposit(shape, X, X, P, N, hw, w, r);                               /* E[1,:] = solve(rho(X,X))  :9 */
    oracle_solve1d(r, N, E1);
    for (int it = 0; it < max_sweeps; ++it) {                                 /* :10 */
        if (oracle_isapprox(F3N, E3N, 3 * N, rtol, atol)) break;
        memcpy(F3N, E3N, sizeof(double) * (size_t)(3 * N));
        for (int64_t j = 0; j < P; ++j) x[j] = X[j] + (v[j] + V[j]) / 2 * dt; /* :11 */
        for (int64_t j = 0; j < P; ++j)                                       /* :12  ...*dt/6 */
            v[j] = V[j] + shape_gather(shape, E1, (X[j] + X[j]) / 2, N, hw) * dt / 6;
        shape_deposit(shape, X, x, P, N, hw, w, r);                           /* :13 */
        oracle_solve1d(r, N, E2);
        for (int64_t j = 0; j < P; ++j)                                       /* :14  ...*4dt/6 */
            v[j] = v[j] + shape_gather(shape, E2, (X[j] + x[j]) / 2, N, hw) * (4 * dt) / 6;
        shape_deposit(shape, x, x, P, N, hw, w, r);                           /* :15 */
        oracle_solve1d(r, N, E3);
        for (int64_t j = 0; j < P; ++j)                                       /* :16 */
            v[j] = v[j] + shape_gather(shape, E3, (x[j] + x[j]) / 2, N, hw) * dt / 6;
        ++sweeps;
    }
    for (int64_t j = 0; j < P; ++j) x[j] = oracle_jl_mod1(x[j]);              /* :17 */
    double se = 0, sv2 = 0, svp = 0;
    for (int64_t i = 0; i < N; ++i) se += E3[i] * E3[i];                      /* sum(E[end,:].^2)  :18 */
    for (int64_t j = 0; j < P; ++j) { sv2 += v[j] * v[j]; svp += v[j] / (double)P; }
    if (D4) {
        double d1 = (se / (double)N) / 2, d2 = (sv2 * W / (double)P) / 2, s = 2 / W;
        D4[0] = d1 * s; D4[1] = d2 * s; D4[2] = (d1 + d2) * s; D4[3] = svp;
    }
    return sweeps;
}

ORACLE_API void oracle_simpson_run(double *x, double *v, double *E3N, int64_t P, int64_t N, int hw, double dt, double W,
                                   double w, double rtol, double atol, int max_sweeps, int64_t T, double *D, int32_t *sweeps,
                                   int shape)
{
    double *X = (double *)malloc(sizeof(double) * (size_t)(2 * P + 4 * N));
    double *V = X + P, *F = V + P, *r = F + 3 * N;
    for (int64_t t = 0; t < T; ++t) {
        double d4[4];
        int s = oracle_simpson_step(x, v, E3N, X, V, F, r, P, N, hw, dt, W, w, rtol, atol, max_sweeps, d4, shape);
        if (D) for (int c = 0; c < 4; ++c) D[c * T + t] = d4[c];
        if (sweeps) sweeps[t] = s;
    }
    free(X);
}

/* Explicit Gaussian leapfrog: Gaussian.jl:8-11.  scale = w/dx (Gaussian.jl:7), hw = 6. */
ORACLE_API void oracle_gauss_leapfrog_step(double *x, double *v, int64_t P, int64_t N, int hw, double dt,
                                           double scale, double *rho, double *E, double *raw)
{
    oracle_ngp_half_drift(x, v, P, dt);                                   /* u() */
    oracle_gauss_deposit(x, x, P, N, hw, scale, rho);                     /* rho(): d.(x); (x+x)/2 == x exactly */
    oracle_solve1d(rho, N, E);
    oracle_ngp_half_drift(x, v, P, dt);                                   /* u() */
    for (int64_t j = 0; j < P; ++j) v[j] += oracle_gauss_gather(E, x[j], N, hw) * dt; /* Gaussian.jl:10 */
    if (raw) {
        double se = 0, sv2 = 0, sv = 0;
        for (int64_t i = 0; i < N; ++i) se += E[i] * E[i];
        for (int64_t j = 0; j < P; ++j) { sv2 += v[j] * v[j]; sv += v[j]; }
        raw[0] = se; raw[1] = sv2; raw[2] = sv;
    }
}

/* boris(vx,vy,E,B,dt) of the 1D2V magnetised codes: src/NGP1D2V.jl:5-10 (B along z). */
ORACLE_API void oracle_boris_1d2v(double *vx, double *vy, double E, double B, double dt)
{
    double m1 = *vx + E * dt / 2, m2 = *vy, m3 = 0.0;            /* v- */
    double t1 = 0.0, t2 = 0.0, t3 = B * dt / 2;
    double c1 = m2 * t3 - m3 * t2, c2 = m3 * t1 - m1 * t3, c3 = m1 * t2 - m2 * t1;
    double p1 = m1 + c1, p2 = m2 + c2, p3 = m3 + c3;
    double q1 = p2 * t3 - p3 * t2, q2 = p3 * t1 - p1 * t3;
    double den = 1 + (t1 * t1 + t2 * t2 + t3 * t3);
    double r1 = m1 + 2 * q1 / den, r2 = m2 + 2 * q2 / den;       /* v+ = v- + 2*cross(...)/(1+dot(t,t)) */
    (void)c3;
    *vx = r1 + E * dt / 2; *vy = r2;
}

/* One step of src/NGP1D2V.jl:40-45,55: E = solve(rho(x)); gather at x, boris, x += vx*dt; x = mod(x,1).
 * rho/E out; raw[0..3] = sum(abs2,E), sum(vx^2+vy^2), sum(vx), sum(vy) after the push (for D, :59-61). */
ORACLE_API void oracle_1d2v_step(double *x, double *vx, double *vy, int64_t P, int64_t N, int hw, double dt, double B0,
                                 double w, double *rho, double *E, double *raw)
{
    oracle_gauss_deposit(x, x, P, N, hw, w, rho);                 /* rho(x): d.((x.+x)./2) */
    oracle_solve1d(rho, N, E);
    for (int64_t j = 0; j < P; ++j) {
        double Ej = oracle_gauss_gather(E, x[j], N, hw);          /* :42 */
        oracle_boris_1d2v(&vx[j], &vy[j], Ej, B0, dt);            /* :43 */
        x[j] += vx[j] * dt;                                       /* :44 */
    }
    for (int64_t j = 0; j < P; ++j) x[j] = oracle_jl_mod1(x[j]); /* :55 */
    if (raw) {
        double se = 0, s0 = 0, s1 = 0, s2 = 0;
        for (int64_t i = 0; i < N; ++i) se += E[i] * E[i];
        for (int64_t j = 0; j < P; ++j) { s0 += vy[j] * vy[j] + vx[j] * vx[j]; s1 += vx[j]; s2 += vy[j]; }
        raw[0] = se; raw[1] = s0; raw[2] = s1; raw[3] = s2;
    }
}

/* boris(vx,vy,E,B,dt,q_m) of the two-species code: src/NGP1D2V2S.jl:5-11 (dt2q_m = dt/2*q_m scales both E and B). */
ORACLE_API void oracle_boris_1d2v_qm(double *vx, double *vy, double E, double B, double dt, double q_m)
{
    double h = dt / 2 * q_m;
    double m1 = *vx + E * h, m2 = *vy, m3 = 0.0;                 /* v- */
    double t1 = 0.0, t2 = 0.0, t3 = B * h;
    double c1 = m2 * t3 - m3 * t2, c2 = m3 * t1 - m1 * t3;
    double p1 = m1 + c1, p2 = m2 + c2, p3 = m3 + (m1 * t2 - m2 * t1);
    double q1 = p2 * t3 - p3 * t2, q2 = p3 * t1 - p1 * t3;
    double den = 1 + (t1 * t1 + t2 * t2 + t3 * t3);
    double r1 = m1 + 2 * q1 / den, r2 = m2 + 2 * q2 / den;
    *vx = r1 + E * h; *vy = r2;
}

/* One step of src/NGP1D2V2S.jl:31-49 on the concatenated arrays [species 1 (q=-1, q/m=-1) | species 2 (q=+1, q/m=1/M)],
 * P particles per species: rho() = rho(x1,-1) then rho(x2,+1) with r[k[1]] += q*k[2]*w (:24-25); solve; per species
 * gather at x, boris(..., q_m), x += vx*dt; x = mod(x,1).  raw[0..3] = sum(abs2,E),
 * sum(vy1^2+vx1^2 + M*(vy2^2+vx2^2)), sum(vx1 + M*vx2), sum(vy1 + M*vy2)  (the sums behind D, :51-52). */
ORACLE_API void oracle_1d2v2s_step(double *x, double *vx, double *vy, int64_t P, int64_t N, int hw, double dt, double B0,
                                   double w, double M, double *rho, double *E, double *raw)
{
    int32_t idx[32]; double wt[32];
    for (int64_t i = 0; i < N; ++i) rho[i] = 0.0;
    for (int sp = 0; sp < 2; ++sp) {
        double q = sp == 0 ? -1.0 : 1.0;
        for (int64_t j = sp * P; j < (sp + 1) * P; ++j) {
            oracle_gauss_stencil(x[j], N, hw, idx, wt);
            for (int k = 0; k <= 2 * hw; ++k) rho[idx[k] - 1] += q * wt[k] * w;
        }
    }
    oracle_solve1d(rho, N, E);
    for (int sp = 0; sp < 2; ++sp) {
        double q_m = sp == 0 ? -1.0 : 1 / M;
        for (int64_t j = sp * P; j < (sp + 1) * P; ++j) {
            double Ej = oracle_gauss_gather(E, x[j], N, hw);
            oracle_boris_1d2v_qm(&vx[j], &vy[j], Ej, B0, dt, q_m);
            x[j] += vx[j] * dt;
        }
    }
    for (int64_t j = 0; j < 2 * P; ++j) x[j] = oracle_jl_mod1(x[j]);
    if (raw) {
        double se = 0, s0 = 0, s1 = 0, s2 = 0;
        for (int64_t i = 0; i < N; ++i) se += E[i] * E[i];
        for (int64_t i = 0; i < P; ++i) {
            s0 += vy[i] * vy[i] + vx[i] * vx[i] + M * (vy[P + i] * vy[P + i] + vx[P + i] * vx[P + i]);
            s1 += vx[i] + M * vx[P + i];
            s2 += vy[i] + M * vy[P + i];
        }
        raw[0] = se; raw[1] = s0; raw[2] = s1; raw[3] = s2;
    }
}

/* Quiet start: x=(bitreverse.(0:P-1).+2.0^63)/2.0^64; v = (j>P/2) ? 1 : -1   GaussianFixedPointQuiet.jl:2-3.
 * Generates global indices [first, first+count) of a P-particle population. */
ORACLE_API void oracle_quiet_start(int64_t P, int64_t first, int64_t count, double *x, double *v)
{
    for (int64_t n = 0; n < count; ++n) {
        uint64_t i = (uint64_t)(first + n), r = 0;
        for (int b = 0; b < 64; ++b) if (i & (1ULL << b)) r |= 1ULL << (63 - b);
        int64_t s = (int64_t)r; /* bitreverse(::Int64) is a signed reinterpretation */
        x[n] = ((double)s + 9223372036854775808.0) / 18446744073709551616.0;
        int64_t j1 = first + n + 1; /* 1-based */
        v[n] = ((double)j1 > (double)P / 2) ? 1.0 : -1.0;
    }
}

/* gamma(x)=imag(sqrt(Complex(x^2+1-sqrt(4x^2+1))))*sqrt(W/2)/log(10)  GaussianFixedPointQuiet.jl:19.
 * Returns the predicted slope 2*gamma(2pi/sqrt(W/2)) of log10 D[:,1] per unit time (line 20). */
ORACLE_API double oracle_growth_slope(double W)
{
    double xk = TWO_PI / sqrt(W / 2);
    double a = xk * xk + 1 - sqrt(4 * xk * xk + 1);
    double im = a < 0 ? sqrt(-a) : 0.0;
    return 2 * im * sqrt(W / 2) / log(10.0);
}

/* ------------------------------------------------------------------ */
/* 2D3V electrostatic CIC + Boris (Electrostatic2D3V.jl)               */
/* ------------------------------------------------------------------ */

/* unimod(x, n) = 0 < x <= n ? x : x > n ? x - n : x + n     Electrostatic2D3V.jl:83 */
static inline double unimod_d(double x, double n) { return (0 < x && x <= n) ? x : (x > n ? x - n : x + n); }
static inline int64_t unimod_i(int64_t x, int64_t n) { return (0 < x && x <= n) ? x : (x > n ? x - n : x + n); }

/* g(z, NZ): Electrostatic2D3V.jl:84-92.  Returns 1-based (i0, w0=1-r), (i1, w1=r). */
static inline void cic_g(double z, int64_t NZ, int64_t *i0, double *w0, int64_t *i1, double *w1)
{
    double zNZ = z * (double)NZ;
    int64_t i = unimod_i((int64_t)ceil(zNZ), NZ);
    double r = (double)i - zNZ;
    *i0 = i; *w0 = 1 - r; *i1 = unimod_i(i + 1, NZ); *w1 = r;
}

ORACLE_API void oracle_cic_g(double z, int64_t NZ, int32_t *idx1, double *wt)
{
    int64_t a, b; cic_g(z, NZ, &a, &wt[0], &b, &wt[1]); idx1[0] = (int32_t)a; idx1[1] = (int32_t)b;
}

/* boris(vx,vy,vz,Ex,Ey,dt): Electrostatic2D3V.jl:32-41; tvec=[B0*dt/2,0,0], tscale=2/(1+dot(tvec,tvec)) */
ORACLE_API void oracle_boris(double *vx, double *vy, double *vz, double Ex, double Ey, double dt, double B0)
{
    double t1 = B0 * dt / 2, t2 = 0.0, t3 = 0.0;
    double tscale = 2 / (1 + (t1 * t1 + t2 * t2 + t3 * t3));
    double dt_2 = dt / 2;
    double e1 = Ex * dt_2, e2 = Ey * dt_2, e3 = 0.0;
    double m1 = *vx + e1, m2 = *vy + e2, m3 = *vz + e3;          /* v- */
    /* cross(a,b) = (a2*b3-a3*b2, a3*b1-a1*b3, a1*b2-a2*b1) */
    double c1 = m2 * t3 - m3 * t2, c2 = m3 * t1 - m1 * t3, c3 = m1 * t2 - m2 * t1;
    double p1 = m1 + c1, p2 = m2 + c2, p3 = m3 + c3;             /* v- + v- x t */
    double q1 = p2 * t3 - p3 * t2, q2 = p3 * t1 - p1 * t3, q3 = p1 * t2 - p2 * t1;
    double r1 = m1 + q1 * tscale, r2 = m2 + q2 * tscale, r3 = m3 + q3 * tscale; /* v+ */
    *vx = r1 + e1; *vy = r2 + e2; *vz = r3 + e3;
}

/* eval(F1,F2,xi,yi): Electrostatic2D3V.jl:94-103 -- outer j over g(y), inner i over g(x), @muladd. */
static inline void cic_eval(const double *F1, const double *F2, int64_t NX, int64_t NY, double xi, double yi,
                            double *o1, double *o2)
{
    int64_t ix[2], iy[2]; double wx[2], wy[2];
    cic_g(xi, NX, &ix[0], &wx[0], &ix[1], &wx[1]);
    cic_g(yi, NY, &iy[0], &wy[0], &iy[1], &wy[1]);
    double a = 0.0, b = 0.0;
    for (int jj = 0; jj < 2; ++jj)
        for (int ii = 0; ii < 2; ++ii) {
            double wxy = wx[ii] * wy[jj];
            int64_t k = (ix[ii] - 1) + (iy[jj] - 1) * NX;
            a = fma(F1[k], wxy, a);
            b = fma(F2[k], wxy, b);
        }
    *o1 = a; *o2 = b;
}

/* depositcharge!(F,x,y,w): F[i,j] += wx*wy*w      Electrostatic2D3V.jl:105-109 */
static inline void cic_deposit(double *F, int64_t NX, int64_t NY, double x, double y, double w)
{
    int64_t ix[2], iy[2]; double wx[2], wy[2];
    cic_g(x, NX, &ix[0], &wx[0], &ix[1], &wx[1]);
    cic_g(y, NY, &iy[0], &wy[0], &iy[1], &wy[1]);
    for (int jj = 0; jj < 2; ++jj)
        for (int ii = 0; ii < 2; ++ii)
            F[(ix[ii] - 1) + (iy[jj] - 1) * NX] += wx[ii] * wy[jj] * w;
}

ORACLE_API void oracle_cic_deposit(const double *x, const double *y, int64_t P, int64_t NX, int64_t NY,
                                   double w, double *rho)
{
    for (int64_t i = 0; i < NX * NY; ++i) rho[i] = 0.0;
    for (int64_t j = 0; j < P; ++j) cic_deposit(rho, NX, NY, x[j], y[j], w);
}

ORACLE_API void oracle_cic_gather(const double *Ex, const double *Ey, int64_t NX, int64_t NY, const double *x,
                                  const double *y, int64_t P, double *ex, double *ey)
{
    for (int64_t j = 0; j < P; ++j) cic_eval(Ex, Ey, NX, NY, x[j], y[j], &ex[j], &ey[j]);
}

static void fft2(double *re, double *im, int64_t NX, int64_t NY, int sign)
{
    /* column-major NX x NY: first along x (contiguous), then along y */
    for (int64_t j = 0; j < NY; ++j) fft1(re + j * NX, im + j * NX, NX, sign);
    double *tr = (double *)malloc(sizeof(double) * 2 * (size_t)NY), *ti = tr + NY;
    for (int64_t i = 0; i < NX; ++i) {
        for (int64_t j = 0; j < NY; ++j) { tr[j] = re[i + j * NX]; ti[j] = im[i + j * NX]; }
        fft1(tr, ti, NY, sign);
        for (int64_t j = 0; j < NY; ++j) { re[i + j * NX] = tr[j]; im[i + j * NX] = ti[j]; }
    }
    free(tr);
}

/* "field invert" + "field solve": Electrostatic2D3V.jl:70-81,142-157.
 * kx=2pi*vcat(0:NX/2-1,-NX/2:-1); minvkk=-im/(kx^2+ky^2), [1,1]=0; fft!(phi); phi[1,1]=0;
 * tmp=phi*minvkk; Ex=tmp*kx[i]; Ey=tmp*ky[j]; ifft!(Ex); ifft!(Ey).  Outputs real parts (what
 * eval and K[ti,1] consume, :99-100,166). */
ORACLE_API void oracle_solve2d(const double *rho, int64_t NX, int64_t NY, double *Ex, double *Ey)
{
    size_t n = (size_t)(NX * NY);
    double *pr = (double *)malloc(sizeof(double) * 6 * n);
    double *pi_ = pr + n, *xr = pi_ + n, *xi = xr + n, *yr = xi + n, *yi = yr + n;
    for (size_t k = 0; k < n; ++k) { pr[k] = rho[k]; pi_[k] = 0.0; }
    fft2(pr, pi_, NX, NY, -1);
    pr[0] = 0.0; pi_[0] = 0.0;
    for (int64_t j = 0; j < NY; ++j) {
        double ky = TWO_PI * (double)(j < NY / 2 ? j : j - NY);
        for (int64_t i = 0; i < NX; ++i) {
            double kx = TWO_PI * (double)(i < NX / 2 ? i : i - NX);
            size_t k = (size_t)(i + j * NX);
            double m = (i == 0 && j == 0) ? 0.0 : -1.0 / (kx * kx + ky * ky); /* minvkk = (0, m) */
            double a = pr[k], b = pi_[k];
            double tre = a * 0.0 - b * m, tim = a * m + b * 0.0;               /* phi*minvkk */
            xr[k] = tre * kx; xi[k] = tim * kx;
            yr[k] = tre * ky; yi[k] = tim * ky;
        }
    }
    fft2(xr, xi, NX, NY, +1);
    fft2(yr, yi, NX, NY, +1);
    double inv = (double)(NX * NY);
    for (size_t k = 0; k < n; ++k) { Ex[k] = xr[k] / inv; Ey[k] = yr[k] / inv; }
    free(pr);
}

struct chunk_job {
    double *x, *y, *vx, *vy, *vz; const double *Ex, *Ey; double *grid;
    int64_t lo, hi, NX, NY; double dt, B0, w;
};

static void *chunk_run(void *arg)
{
    struct chunk_job *c = (struct chunk_job *)arg;
    for (int64_t i = c->lo; i < c->hi; ++i) {
        double exi, eyi;
        cic_eval(c->Ex, c->Ey, c->NX, c->NY, c->x[i], c->y[i], &exi, &eyi);      /* :129 */
        oracle_boris(&c->vx[i], &c->vy[i], &c->vz[i], exi, eyi, c->dt, c->B0);     /* :130 */
        c->x[i] = unimod_d(c->x[i] + c->vx[i] * c->dt, 1);                         /* :131 */
        c->y[i] = unimod_d(c->y[i] + c->vy[i] * c->dt, 1);                         /* :132 */
        cic_deposit(c->grid, c->NX, c->NY, c->x[i], c->y[i], c->w);                /* :135 */
    }
    return NULL;
}

/* One time step of Electrostatic2D3V.jl:120-157.  Ex,Ey in: field from the previous step (zeros at t=1);
 * out: field of this step.  rho out.  nthreads > 1 reproduces the reference's chunked per-thread grids
 * (:114,126-141): chunks of ceil(P/nthreads), phi = sum(ns, dims=3) in thread order. */
ORACLE_API void oracle_2d3v_step(double *x, double *y, double *vx, double *vy, double *vz, int64_t P,
                                 int64_t NX, int64_t NY, double dt, double B0, double w, double *Ex,
                                 double *Ey, double *rho, int nthreads)
{
    size_t n = (size_t)(NX * NY);
    if (nthreads < 1) nthreads = 1;
    double *ns = (double *)calloc(n * (size_t)nthreads, sizeof(double));
    struct chunk_job *jobs = (struct chunk_job *)malloc(sizeof(struct chunk_job) * (size_t)nthreads);
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int64_t chunk = (P + nthreads - 1) / nthreads;
    for (int th = 0; th < nthreads; ++th) {
        int64_t lo = th * chunk, hi = lo + chunk < P ? lo + chunk : P;
        if (lo > P) lo = P;
        struct chunk_job j = { x, y, vx, vy, vz, Ex, Ey, ns + (size_t)th * n, lo, hi, NX, NY, dt, B0, w };
        jobs[th] = j;
        if (nthreads > 1) pthread_create(&tid[th], NULL, chunk_run, &jobs[th]);
        else chunk_run(&jobs[th]);
    }
    if (nthreads > 1) for (int th = 0; th < nthreads; ++th) pthread_join(tid[th], NULL);
    for (size_t k = 0; k < n; ++k) {
        double s = ns[k];
        for (int th = 1; th < nthreads; ++th) s += ns[(size_t)th * n + k];
        rho[k] = s;
    }
    free(ns); free(jobs); free(tid);
    oracle_solve2d(rho, NX, NY, Ex, Ey);
}

/* K[ti,1:5]: Electrostatic2D3V.jl:166-170 */
ORACLE_API void oracle_2d3v_diagnostics(const double *Ex, const double *Ey, int64_t NX, int64_t NY,
                                        const double *vx, const double *vy, int64_t P, double w, double *K5)
{
    double se = 0, sk = 0, sx = 0, sy = 0;
    for (int64_t k = 0; k < NX * NY; ++k) se += Ex[k] * Ex[k] + Ey[k] * Ey[k];
    for (int64_t j = 0; j < P; ++j) { sk += (vx[j] * vx[j] + vy[j] * vy[j]) * w; sx += vx[j]; sy += vy[j]; }
    K5[0] = se / (double)(NX * NY); K5[1] = sk; K5[2] = K5[0] + K5[1]; K5[3] = sx / (double)P; K5[4] = sy / (double)P;
}


/* ------------------------------------------------------------------ */
/* PIC2D3V.jl: the ElectrostaticField path          */
/* Species / shapes / halo grids / loop! / diagnose!, restated literally. */
/* ------------------------------------------------------------------ */

/* unimod(x, n) = x > n ? x - n : x > 0 ? x : x + n        PIC2D3V.jl:10 */
static inline double es_unimod_d(double x, double n) { return x > n ? x - n : (x > 0 ? x : x + n); }
static inline int64_t es_unimod_i(int64_t x, int64_t n) { return x > n ? x - n : (x > 0 ? x : x + n); }

/* halton(i, base, seed)                                   PIC2D3V.jl:29-37 */
ORACLE_API double oracle_es_halton(int64_t i, int64_t base, double seed)
{
    double result = 0.0, f = 1.0;
    while (i > 0) {
        f = f / (double)base;
        result += f * (double)(i % base);
        i /= base;
    }
    return oracle_jl_mod1(result + seed);
}

/* bspline(::BSplineWeighting{N}, x)                       PIC2D3V.jl:1124-1163 (written @fastmath there: the compiler
 * may reassociate, so the last bits are not defined by the source; this is the plain left-to-right reading). */
static int es_bspline(int N, double x, double *f)
{
#define P2(a) ((a) * (a))
#define P3(a) ((a) * (a) * (a))
#define P4(a) (((a) * (a)) * ((a) * (a)))
#define P5(a) ((((a) * (a)) * ((a) * (a))) * (a))
    switch (N) {
    case 0: f[0] = 1.0; return 1;
    case 1: f[0] = x; f[1] = 1 - x; return 2;
    case 2:
        f[0] = 9.0 / 8 + 3.0 / 2 * (x - 1.5) + 1.0 / 2 * P2(x - 1.5);
        f[1] = 3.0 / 4 - P2(x - 0.5);
        f[2] = 9.0 / 8 - 3.0 / 2 * (x + 0.5) + 1.0 / 2 * P2(x + 0.5);
        return 3;
    case 3:
        f[0] = 4.0 / 3 + 2 * (x - 2) + P2(x - 2) + 1.0 / 6 * P3(x - 2);
        f[1] = 2.0 / 3 - P2(x - 1) - 1.0 / 2 * P3(x - 1);
        f[2] = 2.0 / 3 - P2(x) + 1.0 / 2 * P3(x);
        f[3] = 4.0 / 3 - 2 * (x + 1) + P2(x + 1) - 1.0 / 6 * P3(x + 1);
        return 4;
    case 4:
        f[0] = 625.0 / 384 + 125.0 / 48 * (x - 2.5) + 25.0 / 16 * P2(x - 2.5) + 5.0 / 12 * P3(x - 2.5) + 1.0 / 24 * P4(x - 2.5);
        f[1] = 55.0 / 96 - 5.0 / 24 * (x - 1.5) - 5.0 / 4 * P2(x - 1.5) - 5.0 / 6 * P3(x - 1.5) - 1.0 / 6 * P4(x - 1.5);
        f[2] = 115.0 / 192 - 5.0 / 8 * P2(x - 0.5) + 1.0 / 4 * P4(x - 0.5);
        f[3] = 55.0 / 96 + 5.0 / 24 * (x + 0.5) - 5.0 / 4 * P2(x + 0.5) + 5.0 / 6 * P3(x + 0.5) - 1.0 / 6 * P4(x + 0.5);
        f[4] = 625.0 / 384 - 125.0 / 48 * (x + 1.5) + 25.0 / 16 * P2(x + 1.5) - 5.0 / 12 * P3(x + 1.5) + 1.0 / 24 * P4(x + 1.5);
        return 5;
    case 5:
        f[0] = 243.0 / 120 + 81.0 / 24 * (x - 3) + 9.0 / 4 * P2(x - 3) + 3.0 / 4 * P3(x - 3) + 1.0 / 8 * P4(x - 3) + 1.0 / 120 * P5(x - 3);
        f[1] = 17.0 / 40 - 5.0 / 8 * (x - 2) - 7.0 / 4 * P2(x - 2) - 5.0 / 4 * P3(x - 2) - 3.0 / 8 * P4(x - 2) - 1.0 / 24 * P5(x - 2);
        f[2] = 22.0 / 40 - 1.0 / 2 * P2(x - 1) + 1.0 / 4 * P4(x - 1) + 1.0 / 12 * P5(x - 1);
        f[3] = 22.0 / 40 - 1.0 / 2 * P2(x + 0) + 1.0 / 4 * P4(x - 0) - 1.0 / 12 * P5(x - 0);
        f[4] = 17.0 / 40 + 5.0 / 8 * (x + 1) - 7.0 / 4 * P2(x + 1) + 5.0 / 4 * P3(x + 1) - 3.0 / 8 * P4(x + 1) + 1.0 / 24 * P5(x + 1);
        f[5] = 243.0 / 120 - 81.0 / 24 * (x + 2) + 9.0 / 4 * P2(x + 2) - 3.0 / 4 * P3(x + 2) + 1.0 / 8 * P4(x + 2) - 1.0 / 120 * P5(x + 2);
        return 6;
    }
#undef P2
#undef P3
#undef P4
#undef P5
    return 0;
}

/* depositindicesfractions(s, z, NZ, NZ_Lz) + gridinteractiontuple   PIC2D3V.jl:1105-1121,1165-1188.
 * shape: 0 NGPWeighting, 1 AreaWeighting, 10+N BSplineWeighting{N}.  Returns the number of (index, fraction) pairs;
 * *j0 is the first index, 1-based and NOT wrapped ("no need for unimod with offset arrays" :1108,1117).
 * The reference asserts 0 < r <= 1 (:1111); r == 0 (z exactly on a cell edge) throws there and is carried through here. */
static int es_shape(int shape, double z, double NZ_Lz, int64_t *j0, double *wt)
{
    double zNZ = z * NZ_Lz;
    int64_t i = (int64_t)ceil(zNZ);
    double r = (double)i - zNZ;
    if (shape == 0) { *j0 = i; wt[0] = 1; return 1; }                       /* ((i, 1), ) */
    if (shape == 1) { *j0 = i; wt[0] = 1 - r; wt[1] = r; return 2; }        /* ((i, 1-r), (i+1, r)) */
    int N = shape - 10;
    int64_t j; double zz;
    if (N & 1) { j = i; zz = 1 - r; }                                       /* _bsplineinputs, odd N  :1168 */
    else { int q = r > 0.5; j = i + q; zz = (double)q + 0.5 - r; }          /* even N                 :1169-1172 */
    *j0 = j - N / 2;                                                        /* indices: (j-fld(N,2)):(j+cld(N,2)) */
    return es_bspline(N, zz, wt);
}

ORACLE_API int oracle_es_shape(int shape, double z, double NZ_Lz, int64_t *j0, double *wt6)
{
    return es_shape(shape, z, NZ_Lz, j0, wt6);
}

/* (boris::ElectrostaticBoris)(vx, vy, vz, Ex, Ey, q_m)     PIC2D3V.jl:44-54; t = B*dt/2, t2 = dot(t,t). */
static inline void es_cross(const double *a, const double *b, double *c)
{
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
ORACLE_API void oracle_es_boris(double *v, double Ex, double Ey, const double *B, double dt, double q_m)
{
    double t[3] = { B[0] * dt / 2, B[1] * dt / 2, B[2] * dt / 2 };
    double t2 = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
    double dt_2 = dt / 2;
    double e2[3] = { Ex * dt_2 * q_m, Ey * dt_2 * q_m, 0.0 * dt_2 * q_m };
    double vm[3] = { v[0] + e2[0], v[1] + e2[1], v[2] + e2[2] };
    double c1[3], s[3], c2[3];
    es_cross(vm, t, c1);
    for (int k = 0; k < 3; ++k) s[k] = vm[k] + c1[k];
    es_cross(s, t, c2);
    double den = 1 + q_m * q_m * t2;
    for (int k = 0; k < 3; ++k) v[k] = (vm[k] + c2[k] * (q_m * q_m) * 2 / den) + e2[k];
}

/* Halo ("offset") arrays of ElectrostaticField, buffer = 3: indices -(buffer-1):NZ+buffer   PIC2D3V.jl:284-287 */
#define ES_BUF 3
static inline size_t es_hidx(int64_t i, int64_t j, int64_t NX) { return (size_t)((i + ES_BUF - 1) + (NX + 2 * ES_BUF) * (j + ES_BUF - 1)); }

/* One call of loop!(plasma, field::ElectrostaticField, ...)   PIC2D3V.jl:530-581, followed by update! :294-297.
 * Particles: species one after the other in x,y,vx,vy,vz (SoA image of xyv[5,P]).
 * Exy: [2][(NX+6)*(NY+6)] halo arrays, in/out.  accumulate = 1 is the reference as written: update! ADDS the new
 * field to Exy (applyperiodicity!(oa, a) does oa[i,j] += real(a[...]) :21-27 and nothing zeroes Exy for this field
 * type); accumulate = 0 zeroes Exy first (what the Lorenz-gauge update! does, :475).
 * Outputs (NX*NY, column-major): rho = reduction!(phi, rhos) before the transform; Ex, Ey = real parts after the inverse
 * transforms; phir = real(pifft! * phi) of :1326 (phi holds the spectrum of rho with [1,1] zeroed, i.e. rho - mean). */
ORACLE_API void oracle_es_loop(int nspecies, const int64_t *sP, const int32_t *sshape, const double *scharge,
                               const double *smass, const double *sweight, double *x, double *y, double *vx, double *vy,
                               double *vz, int64_t NX, int64_t NY, double Lx, double Ly, double dt, const double *B,
                               double *Exy, int accumulate, double *rho, double *Ex, double *Ey, double *phir, int nthreads)
{
    const size_t n = (size_t)(NX * NY), hn = (size_t)((NX + 2 * ES_BUF) * (NY + 2 * ES_BUF));
    const double NX_Lx = (double)NX / Lx, NY_Ly = (double)NY / Ly;
    const double dV = (Lx / (double)NX) * (Ly / (double)NY);
    if (nthreads < 1) nthreads = 1;
    double *rhos = (double *)calloc(hn * (size_t)nthreads, sizeof(double));
    const double *Fx = Exy, *Fy = Exy + hn;
    for (int k = 0; k < nthreads; ++k) {                                    /* @threads for k in axes(field.rhos, 3) */
        double *rk = rhos + (size_t)k * hn;
        int64_t base = 0;
        for (int s = 0; s < nspecies; ++s) {                                /* for species in plasma */
            const int64_t P = sP[s];
            const double qw_dV = scharge[s] * sweight[s] / dV, q_m = scharge[s] / smass[s];
            const int64_t chunk = (P + nthreads - 1) / nthreads;            /* partition(1:P, ceil(Int, P/nthreads())) */
            int64_t lo = k * chunk, hi = lo + chunk < P ? lo + chunk : P;
            for (int64_t i = base + lo; i < base + hi; ++i) {
                int64_t ix0, iy0; double wx[6], wy[6];
                int nx = es_shape(sshape[s], x[i], NX_Lx, &ix0, wx), ny = es_shape(sshape[s], y[i], NY_Ly, &iy0, wy);
                double Exi = 0, Eyi = 0;                                     /* field(shape, x, y) :1217-1229 */
                for (int b = 0; b < ny; ++b)
                    for (int a = 0; a < nx; ++a) {
                        double wxy = wx[a] * wy[b];
                        Exi = fma(Fx[es_hidx(ix0 + a, iy0 + b, NX)], wxy, Exi);
                        Eyi = fma(Fy[es_hidx(ix0 + a, iy0 + b, NX)], wxy, Eyi);
                    }
                double vxi = vx[i], vyi = vy[i], v[3] = { vx[i], vy[i], vz[i] };
                oracle_es_boris(v, Exi, Eyi, B, dt, q_m);
                vx[i] = v[0]; vy[i] = v[1]; vz[i] = v[2];
                x[i] = es_unimod_d(x[i] + (vxi + vx[i]) / 2 * dt, Lx);
                y[i] = es_unimod_d(y[i] + (vyi + vy[i]) / 2 * dt, Ly);
                nx = es_shape(sshape[s], x[i], NX_Lx, &ix0, wx); ny = es_shape(sshape[s], y[i], NY_Ly, &iy0, wy);
                for (int b = 0; b < ny; ++b)                                 /* deposit! :1246-1253 */
                    for (int a = 0; a < nx; ++a) rk[es_hidx(ix0 + a, iy0 + b, NX)] += wx[a] * wy[b] * qw_dV;
            }
            base += P;
        }
    }
    /* reduction!(field.phi, field.rhos): phi = 0, fold every thread's halo grid with unimod   :485-490, 13-19 */
    double *pr = (double *)calloc(6 * n, sizeof(double));
    double *pim = pr + n, *xr = pim + n, *xi = xr + n, *yr = xi + n, *yi = yr + n;
    for (int k = 0; k < nthreads; ++k)
        for (int64_t j = -(ES_BUF - 1); j <= NY + ES_BUF; ++j)
            for (int64_t i = -(ES_BUF - 1); i <= NX + ES_BUF; ++i)
                pr[(es_unimod_i(i, NX) - 1) + NX * (es_unimod_i(j, NY) - 1)] += rhos[(size_t)k * hn + es_hidx(i, j, NX)];
    free(rhos);
    for (size_t k = 0; k < n; ++k) rho[k] = pr[k];
    fft2(pr, pim, NX, NY, -1);
    pr[0] = 0.0; pim[0] = 0.0;                                               /* field.phi[1, 1] = 0 */
    for (int64_t j = 0; j < NY; ++j) {
        double ky = TWO_PI / Ly * (double)(j < NY / 2 ? j : j - NY);         /* FFTHelper :254-255 */
        for (int64_t i = 0; i < NX; ++i) {
            double kx = TWO_PI / Lx * (double)(i < NX / 2 ? i : i - NX);
            size_t k = (size_t)(i + j * NX);
            double m = (i == 0 && j == 0) ? 0.0 : -1.0 / (kx * kx + ky * ky); /* im_k^-2 = -im ./ k2, [1,1] = 0 */
            double a = pr[k], b = pim[k];
            double tre = a * 0.0 - b * m, tim = a * m + b * 0.0;
            xr[k] = tre * kx; xi[k] = tim * kx; yr[k] = tre * ky; yi[k] = tim * ky;
        }
    }
    fft2(xr, xi, NX, NY, +1);
    fft2(yr, yi, NX, NY, +1);
    fft2(pr, pim, NX, NY, +1);                                               /* diagnose!: real.(pifft! * phi) :1326 */
    const double inv = (double)(NX * NY);
    for (size_t k = 0; k < n; ++k) { Ex[k] = xr[k] / inv; Ey[k] = yr[k] / inv; phir[k] = pr[k] / inv; }
    free(pr);
    /* update!(field): applyperiodicity!(view(Exy,c,:,:), E) -> oa[i,j] += real(a[unimod(i,NX), unimod(j,NY)]) */
    double *Gx = Exy, *Gy = Exy + hn;
    if (!accumulate) memset(Exy, 0, 2 * hn * sizeof(double));
    for (int64_t j = -(ES_BUF - 1); j <= NY + ES_BUF; ++j)
        for (int64_t i = -(ES_BUF - 1); i <= NX + ES_BUF; ++i) {
            size_t src = (size_t)((es_unimod_i(i, NX) - 1) + NX * (es_unimod_i(j, NY) - 1));
            Gx[es_hidx(i, j, NX)] += Ex[src];
            Gy[es_hidx(i, j, NX)] += Ey[src];
        }
}

/* diagnose!(d, plasma) + the field energy of diagnose!(d::ElectrostaticDiagnostics, f, ...)   PIC2D3V.jl:1301-1321.
 * out[8] = kineticenergy, fieldenergy, particlemomentum[3], characteristicmomentum[3].
 * kineticenergy(s) = sum(abs2, velocities(s)) * s.mass / 2 * s.weight  :177; momentum(s, op) :179-187;
 * fieldenergy = mean(abs2, f.Exy) / 2 over the WHOLE halo array (2 x (NX+6) x (NY+6) elements). */
ORACLE_API void oracle_es_diagnose(int nspecies, const int64_t *sP, const double *smass, const double *sweight,
                                   const double *vx, const double *vy, const double *vz, int64_t NX, int64_t NY,
                                   const double *Exy, double *out)
{
    for (int k = 0; k < 8; ++k) out[k] = 0.0;
    int64_t base = 0;
    for (int s = 0; s < nspecies; ++s) {
        double s2 = 0, m[3] = { 0, 0, 0 }, c[3] = { 0, 0, 0 };
        for (int64_t i = base; i < base + sP[s]; ++i) {
            s2 += vx[i] * vx[i]; s2 += vy[i] * vy[i]; s2 += vz[i] * vz[i];
            m[0] += vx[i]; m[1] += vy[i]; m[2] += vz[i];
            c[0] += fabs(vx[i]); c[1] += fabs(vy[i]); c[2] += fabs(vz[i]);
        }
        out[0] += s2 * smass[s] / 2 * sweight[s];
        for (int k = 0; k < 3; ++k) { out[2 + k] += m[k] * (smass[s] * sweight[s]); out[5 + k] += c[k] * (smass[s] * sweight[s]); }
        base += sP[s];
    }
    const size_t hn = (size_t)((NX + 2 * ES_BUF) * (NY + 2 * ES_BUF));
    double se = 0;
    for (size_t k = 0; k < 2 * hn; ++k) se += Exy[k] * Exy[k];
    out[1] = se / (double)(2 * hn) / 2;
}

/* Species(P, vth, density, shape; Lx, Ly)   PIC2D3V.jl:194-213: Halton starts (sample(P,i) = halton.(0:P-1, i, 1/sqrt(2))
 * :191), v = vth * erfinv(2 sample - 1) * vth, mean removed, rescaled to std vth/sqrt(2) (Statistics.std: corrected,
 * n-1).  erfinv is supplied by the caller's table einv[3*P] = erfinv.(2 .* sample(P, {5,7,9}) .- 1) (scipy in the tests):
 * the C library has no erfinv.  Returns the weight calculateweight(n0, P, Lx, Ly) = n0*Lx*Ly/P  :189. */
ORACLE_API double oracle_es_species(int64_t P, double vth, double density, double Lx, double Ly, const double *einv,
                                    double *x, double *y, double *vx, double *vy, double *vz)
{
    const double seed = 1 / sqrt(2.0);
    double *v[3] = { vx, vy, vz };
    for (int64_t i = 0; i < P; ++i) { x[i] = Lx * oracle_es_halton(i, 2, seed); y[i] = Ly * oracle_es_halton(i, 3, seed); }
    for (int c = 0; c < 3; ++c) {
        double *u = v[c];
        for (int64_t i = 0; i < P; ++i) u[i] = vth * einv[(size_t)c * P + i] * vth;
        double mean = 0;
        for (int64_t i = 0; i < P; ++i) mean += u[i];
        mean /= (double)P;
        for (int64_t i = 0; i < P; ++i) u[i] -= mean;
        double m2 = 0, ss = 0;
        for (int64_t i = 0; i < P; ++i) m2 += u[i];
        m2 /= (double)P;
        for (int64_t i = 0; i < P; ++i) ss += (u[i] - m2) * (u[i] - m2);
        double sd = sqrt(ss / (double)(P - 1));
        for (int64_t i = 0; i < P; ++i) u[i] *= (vth / sqrt(2.0)) / sd;
    }
    return density * Lx * Ly / (double)P;
}

ORACLE_API int oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

ORACLE_API int oracle_version(void) { return 1; }
