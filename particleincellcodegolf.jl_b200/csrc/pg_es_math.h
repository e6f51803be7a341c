// pg_es_math.h -- per-particle arithmetic of the PIC2D3V.jl electrostatic path:
//   unimod                      src/PIC2D3V.jl:10
//   depositindicesfractions     src/PIC2D3V.jl:1105-1113  (+ gridinteractiontuple :1115-1121,1176-1188)
//   bspline(BSplineWeighting{N}) src/PIC2D3V.jl:1124-1163, _bsplineinputs / indices :1165-1174
//   ElectrostaticBoris          src/PIC2D3V.jl:39-54
// Plain C++ without CUDA intrinsics: the kernels (pg_esfield.cuh, compiled with -fmad=false) use these functions on
// the device, and tests/test_esfield_cpu.py compiles the very same header with g++ -ffp-contract=off to compare the
// arithmetic with the oracle on the CPU (test infrastructure: the library itself exports no CPU path).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PG_HD __host__ __device__ __forceinline__
#else
#define PG_HD static inline
#endif

namespace pg {
namespace es {

// picgolf_es_shape (include/picgolf_es.h)
enum { SHAPE_NGP = 0, SHAPE_AREA = 1, SHAPE_BSPLINE0 = 10, SHAPE_BSPLINE5 = 15 };

// number of grid points a shape touches per dimension
PG_HD constexpr int support(int shape) { return shape == SHAPE_NGP ? 1 : shape == SHAPE_AREA ? 2 : shape - SHAPE_BSPLINE0 + 1; }
PG_HD bool shape_valid(int shape) { return shape == SHAPE_NGP || shape == SHAPE_AREA || (shape >= SHAPE_BSPLINE0 && shape <= SHAPE_BSPLINE5); }

// unimod(x, n) = x > n ? x - n : x > 0 ? x : x + n      :10
PG_HD double unimod(double x, double n) { return x > n ? x - n : (x > 0 ? x : x + n); }
PG_HD int unimod(int x, int n) { return x > n ? x - n : (x > 0 ? x : x + n); }

PG_HD double p2(double a) { return a * a; }
PG_HD double p3(double a) { return a * a * a; }
PG_HD double p4(double a) { return (a * a) * (a * a); }
PG_HD double p5(double a) { return ((a * a) * (a * a)) * a; }

// bspline(::BSplineWeighting{N}, x): the N+1 fractions, in the reference's order.  The reference evaluates these under
// @fastmath (association is the compiler's choice); this is the left-to-right reading, the same one the oracle uses.
template <int N>
PG_HD void bspline(double x, double *f)
{
    if constexpr (N == 0) { f[0] = 1.0; }
    if constexpr (N == 1) { f[0] = x; f[1] = 1 - x; }
    if constexpr (N == 2) {
        f[0] = 9.0 / 8 + 3.0 / 2 * (x - 1.5) + 1.0 / 2 * p2(x - 1.5);
        f[1] = 3.0 / 4 - p2(x - 0.5);
        f[2] = 9.0 / 8 - 3.0 / 2 * (x + 0.5) + 1.0 / 2 * p2(x + 0.5);
    }
    if constexpr (N == 3) {
        f[0] = 4.0 / 3 + 2 * (x - 2) + p2(x - 2) + 1.0 / 6 * p3(x - 2);
        f[1] = 2.0 / 3 - p2(x - 1) - 1.0 / 2 * p3(x - 1);
        f[2] = 2.0 / 3 - p2(x) + 1.0 / 2 * p3(x);
        f[3] = 4.0 / 3 - 2 * (x + 1) + p2(x + 1) - 1.0 / 6 * p3(x + 1);
    }
    if constexpr (N == 4) {
        f[0] = 625.0 / 384 + 125.0 / 48 * (x - 2.5) + 25.0 / 16 * p2(x - 2.5) + 5.0 / 12 * p3(x - 2.5) + 1.0 / 24 * p4(x - 2.5);
        f[1] = 55.0 / 96 - 5.0 / 24 * (x - 1.5) - 5.0 / 4 * p2(x - 1.5) - 5.0 / 6 * p3(x - 1.5) - 1.0 / 6 * p4(x - 1.5);
        f[2] = 115.0 / 192 - 5.0 / 8 * p2(x - 0.5) + 1.0 / 4 * p4(x - 0.5);
        f[3] = 55.0 / 96 + 5.0 / 24 * (x + 0.5) - 5.0 / 4 * p2(x + 0.5) + 5.0 / 6 * p3(x + 0.5) - 1.0 / 6 * p4(x + 0.5);
        f[4] = 625.0 / 384 - 125.0 / 48 * (x + 1.5) + 25.0 / 16 * p2(x + 1.5) - 5.0 / 12 * p3(x + 1.5) + 1.0 / 24 * p4(x + 1.5);
    }
    if constexpr (N == 5) {
        f[0] = 243.0 / 120 + 81.0 / 24 * (x - 3) + 9.0 / 4 * p2(x - 3) + 3.0 / 4 * p3(x - 3) + 1.0 / 8 * p4(x - 3) + 1.0 / 120 * p5(x - 3);
        f[1] = 17.0 / 40 - 5.0 / 8 * (x - 2) - 7.0 / 4 * p2(x - 2) - 5.0 / 4 * p3(x - 2) - 3.0 / 8 * p4(x - 2) - 1.0 / 24 * p5(x - 2);
        f[2] = 22.0 / 40 - 1.0 / 2 * p2(x - 1) + 1.0 / 4 * p4(x - 1) + 1.0 / 12 * p5(x - 1);
        f[3] = 22.0 / 40 - 1.0 / 2 * p2(x + 0) + 1.0 / 4 * p4(x - 0) - 1.0 / 12 * p5(x - 0);
        f[4] = 17.0 / 40 + 5.0 / 8 * (x + 1) - 7.0 / 4 * p2(x + 1) + 5.0 / 4 * p3(x + 1) - 3.0 / 8 * p4(x + 1) + 1.0 / 24 * p5(x + 1);
        f[5] = 243.0 / 120 - 81.0 / 24 * (x + 2) + 9.0 / 4 * p2(x + 2) - 3.0 / 4 * p3(x + 2) + 1.0 / 8 * p4(x + 2) - 1.0 / 120 * p5(x + 2);
    }
}

// depositindicesfractions(shape, z, NZ, NZ_Lz): first grid index j0 (1-based, NOT wrapped: the reference indexes halo
// "offset" arrays) and the support(SHAPE) fractions.  The reference's `@assert 0 < r <= 1` (:1111) is not raised here:
// a particle exactly on a cell edge (r == 0) gets the weights the formulas give.
template <int SHAPE>
PG_HD void shape_weights(double z, double NZ_Lz, int &j0, double *w)
{
    const double zNZ = z * NZ_Lz;        // position in units of cells
    const int i = (int)ceil(zNZ);        // cell number
    const double r = (double)i - zNZ;    // distance into cell i
    if (SHAPE == SHAPE_NGP) { j0 = i; w[0] = 1.0; return; }                  // ((i, 1), )
    if (SHAPE == SHAPE_AREA) { j0 = i; w[0] = 1 - r; w[1] = r; return; }      // ((i, 1-r), (i+1, r))
    constexpr int N = SHAPE >= SHAPE_BSPLINE0 ? SHAPE - SHAPE_BSPLINE0 : 0;
    int j; double zz;
    if (N & 1) { j = i; zz = 1 - r; }                                         // odd N:  (i, 1 - centre)
    else { const int q = r > 0.5 ? 1 : 0; j = i + q; zz = (double)q + 0.5 - r; } // even N: (i + q, q + 0.5 - centre)
    j0 = j - N / 2;                                                           // (j - fld(N,2)) : (j + cld(N,2))
    bspline<N>(zz, w);
}

// Run-time dispatch (stage entry points and the CPU harness).  Returns the support, 0 for an unknown shape.
PG_HD int shape_weights_rt(int shape, double z, double NZ_Lz, int &j0, double *w)
{
    switch (shape) {
    case SHAPE_NGP: shape_weights<SHAPE_NGP>(z, NZ_Lz, j0, w); return 1;
    case SHAPE_AREA: shape_weights<SHAPE_AREA>(z, NZ_Lz, j0, w); return 2;
    case 10: shape_weights<10>(z, NZ_Lz, j0, w); return 1;
    case 11: shape_weights<11>(z, NZ_Lz, j0, w); return 2;
    case 12: shape_weights<12>(z, NZ_Lz, j0, w); return 3;
    case 13: shape_weights<13>(z, NZ_Lz, j0, w); return 4;
    case 14: shape_weights<14>(z, NZ_Lz, j0, w); return 5;
    case 15: shape_weights<15>(z, NZ_Lz, j0, w); return 6;
    }
    return 0;
}

// ElectrostaticBoris: t = B*dt/2, t2 = dot(t,t), dt_2 = dt/2 (:44-48); the push (:49-54) as written, including its
// use of q_m (E impulse scaled by q_m, rotation about the UNSCALED t with the factor q_m^2*2/(1+q_m^2*t2)).
// The three divisions `... * 2 / (1 + q_m^2 t2)` of a push have the same divisor for every particle of a species, so the
// species' Boris carries den = 1 + q_m^2 t2 and rden = RN(1/den), and the quotient is formed as
//   q0 = x*rden;  e = fma(-den, q0, x);  q = fma(e, rden, q0)
// which is the correctly rounded x/den (Markstein's theorem: q0 is within an ulp of the quotient, e is exact, and rden is
// the correctly rounded reciprocal) -- the same bits as the IEEE division the reference performs, for a third of the
// instructions of a double-precision divide.  tests/test_esfield_cpu.py checks it bit for bit against the oracle's `/`.
struct Boris {
    double t[3], t2, dt_2;
    double q_m, q2, den, rden; // of the species this push is for
};
PG_HD Boris make_boris(double B0x, double B0y, double B0z, double dt, double q_m)
{
    Boris b;
    b.t[0] = B0x * dt / 2; b.t[1] = B0y * dt / 2; b.t[2] = B0z * dt / 2;
    b.t2 = b.t[0] * b.t[0] + b.t[1] * b.t[1] + b.t[2] * b.t[2];
    b.dt_2 = dt / 2;
    b.q_m = q_m; b.q2 = q_m * q_m; b.den = 1 + b.q2 * b.t2; b.rden = 1 / b.den;
    return b;
}
PG_HD double div_by_den(double x, const Boris &b)
{
    const double q0 = x * b.rden;
    const double e = fma(-b.den, q0, x);
    return fma(e, b.rden, q0);
}
PG_HD void boris_push(const Boris &b, double &vx, double &vy, double &vz, double Ex, double Ey)
{
    const double q_m = b.q_m;
    const double e0 = Ex * b.dt_2 * q_m, e1 = Ey * b.dt_2 * q_m, e2 = 0.0 * b.dt_2 * q_m; // E2 = [Ex, Ey, 0.0] * dt_2 * q_m
    const double m0 = vx + e0, m1 = vy + e1, m2 = vz + e2;                                 // v- = v + E2
    // cross(a, b) = (a2*b3 - a3*b2, a3*b1 - a1*b3, a1*b2 - a2*b1)
    const double s0 = m0 + (m1 * b.t[2] - m2 * b.t[1]);
    const double s1 = m1 + (m2 * b.t[0] - m0 * b.t[2]);
    const double s2 = m2 + (m0 * b.t[1] - m1 * b.t[0]);                                    // v- + cross(v-, t)
    const double c0 = s1 * b.t[2] - s2 * b.t[1], c1 = s2 * b.t[0] - s0 * b.t[2], c2 = s0 * b.t[1] - s1 * b.t[0];
    const double q2 = b.q2;                                                                // den = 1 + q2 * t2
    vx = (m0 + div_by_den(c0 * q2 * 2, b)) + e0;                                           // v+ + E2
    vy = (m1 + div_by_den(c1 * q2 * 2, b)) + e1;
    vz = (m2 + div_by_den(c2 * q2 * 2, b)) + e2;
}

// halton(i, base, seed) (:29-37) and sample(P, i) = halton.(0:P-1, i, 1/sqrt(2)) (:191): the quiet start of Species(...)
PG_HD double halton(long long i, int base, double seed)
{
    double result = 0.0, f = 1.0;
    while (i > 0) {
        f = f / (double)base;
        result += f * (double)(i % base);
        i /= base;
    }
    const double a = result + seed, r = a - trunc(a); // Julia mod(a, 1) for a >= 0
    return r;
}

} // namespace es
} // namespace pg
