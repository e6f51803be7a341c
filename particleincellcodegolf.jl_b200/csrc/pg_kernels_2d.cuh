// pg_kernels_2d.cuh -- the fused particle loop of src/Electrostatic2D3V.jl:125-138:
//   gather E at the old position (eval :94-103) -> boris (:35-41) -> move + unimod (:131-132)
//   -> CIC deposit at the new position (depositcharge! :105-109).
// Five fp64 SoA streams in, five out (80 B per particle-step).  The field is kept as one
// (Ex,Ey) double2 per cell so a CIC corner is a single 16-byte load.
#pragma once
#include "pg_common.cuh"

namespace pg {

// g(z,NZ): i=unimod(ceil(Int,z*NZ),NZ); r=i-z*NZ; ((i,1-r),(unimod(i+1,NZ),r))   :84-92.  1-based cells.
__device__ __forceinline__ void cic_g(double z, int NZ, int &i0, double &w0, int &i1, double &w1)
{
    double zNZ = z * (double)NZ;
    int i = unimod((int)ceil(zNZ), NZ);
    double r = (double)i - zNZ;
    i0 = i; w0 = 1 - r; i1 = unimod(i + 1, NZ); w1 = r;
}

// boris(vx,vy,vz,Ex,Ey,dt) with tvec=[B0*dt/2,0,0], tscale=2/(1+dot(tvec,tvec))   :32-41.
// The zero components of tvec are kept symbolically (x*0 terms dropped: they are exact zeros).
__device__ __forceinline__ void boris(double &vx, double &vy, double &vz, double Ex, double Ey, double dt, double t1,
                                      double tscale)
{
    double dt_2 = dt / 2;
    double e1 = Ex * dt_2, e2 = Ey * dt_2;
    double m1 = vx + e1, m2 = vy + e2, m3 = vz;          // v- = v + Edt_2
    // cross(a, t) with t = (t1,0,0) is (0, a3*t1, -(a2*t1)); the a*0 terms are exact zeros.
    double p2 = m2 + m3 * t1, p3 = m3 - m2 * t1;         // v- + v- x t
    double r2 = m2 + (p3 * t1) * tscale, r3 = m3 - (p2 * t1) * tscale; // v+
    vx = m1 + e1; vy = r2 + e2; vz = r3;
}

struct P2DArgs {
    double *x, *y, *vx, *vy, *vz;
    const double2 *E2;   // (real(Ex), real(Ey)) per cell, column-major NX x NY
    double *rho;         // global deposit grid (fp64 atomics)
    double *partials;    // [3*gridDim.x] per-block (sum vx^2+vy^2, sum vx, sum vy)
    long long P;
    double dt, w, t1, tscale;
    int NX, NY;
};

__global__ void __launch_bounds__(PG_THREADS) particles_2d3v_kernel(P2DArgs a)
{
    __shared__ double scratch[32];
    const int NX = a.NX, NY = a.NY;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.P; p += stride) {
        double x = ld_stream(a.x + p), y = ld_stream(a.y + p);
        double vx = ld_stream(a.vx + p), vy = ld_stream(a.vy + p), vz = ld_stream(a.vz + p);
        int ix[2], iy[2];
        double wx[2], wy[2];
        cic_g(x, NX, ix[0], wx[0], ix[1], wx[1]);
        cic_g(y, NY, iy[0], wy[0], iy[1], wy[1]);
        double ex = 0.0, ey = 0.0;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
                double wxy = wx[ii] * wy[jj];
                double2 f = __ldg(&a.E2[(ix[ii] - 1) + (size_t)(iy[jj] - 1) * NX]);
                ex = fma(f.x, wxy, ex); // @muladd F1o += real(F1[i,j]) * wxy
                ey = fma(f.y, wxy, ey);
            }
        boris(vx, vy, vz, ex, ey, a.dt, a.t1, a.tscale);
        x = unimod(x + vx * a.dt, 1.0);
        y = unimod(y + vy * a.dt, 1.0);
        cic_g(x, NX, ix[0], wx[0], ix[1], wx[1]);
        cic_g(y, NY, iy[0], wy[0], iy[1], wy[1]);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
                atomicAdd(&a.rho[(ix[ii] - 1) + (size_t)(iy[jj] - 1) * NX], wx[ii] * wy[jj] * a.w); // F[i,j] += wx*wy*w
        st_stream(a.x + p, x); st_stream(a.y + p, y);
        st_stream(a.vx + p, vx); st_stream(a.vy + p, vy); st_stream(a.vz + p, vz);
        s0 += vx * vx + vy * vy; s1 += vx; s2 += vy;
    }
    s0 = block_sum(s0, scratch);
    s1 = block_sum(s1, scratch);
    s2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) {
        a.partials[3 * blockIdx.x] = s0; a.partials[3 * blockIdx.x + 1] = s1; a.partials[3 * blockIdx.x + 2] = s2;
    }
}

// x,y ~ U(0,1]; v Maxwellian (Box-Muller on splitmix64 draws) with per-component std vth/sqrt(2),
// the distribution Electrostatic2D3V.jl:45-55 prepares (without its sample-mean/std correction).
__global__ void synthetic_2d3v_kernel(double *x, double *y, double *vx, double *vy, double *vz, long long count,
                                      long long first, uint64_t seed, double vth)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double sd = vth / sqrt(2.0);
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < count; n += stride) {
        uint64_t g = (uint64_t)(first + n) * 8ULL;
        double u[6];
        for (int q = 0; q < 6; ++q) u[q] = u01(splitmix64(seed ^ ((g + q) * 0xD1342543DE82EF95ULL)));
        x[n] = 1.0 - u[0]; y[n] = 1.0 - u[1]; // (0,1]
        double r1 = sqrt(-2.0 * log(1.0 - u[2])), r2 = sqrt(-2.0 * log(1.0 - u[4]));
        vx[n] = sd * r1 * cospi(2.0 * u[3]);
        vy[n] = sd * r1 * sinpi(2.0 * u[3]);
        vz[n] = sd * r2 * cospi(2.0 * u[5]);
    }
}

// ---- stage kernels -------------------------------------------------------------------------
__global__ void stage_cic_deposit_kernel(const double *x, const double *y, long long count, int NX, int NY, double w,
                                         double *rho)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    int ix[2], iy[2]; double wx[2], wy[2];
    cic_g(x[p], NX, ix[0], wx[0], ix[1], wx[1]);
    cic_g(y[p], NY, iy[0], wy[0], iy[1], wy[1]);
    for (int jj = 0; jj < 2; ++jj)
        for (int ii = 0; ii < 2; ++ii)
            atomicAdd(&rho[(ix[ii] - 1) + (size_t)(iy[jj] - 1) * NX], wx[ii] * wy[jj] * w);
}

__global__ void stage_cic_gather_kernel(const double *Ex, const double *Ey, int NX, int NY, const double *x,
                                        const double *y, long long count, double *ex, double *ey)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    int ix[2], iy[2]; double wx[2], wy[2];
    cic_g(x[p], NX, ix[0], wx[0], ix[1], wx[1]);
    cic_g(y[p], NY, iy[0], wy[0], iy[1], wy[1]);
    double a = 0.0, b = 0.0;
    for (int jj = 0; jj < 2; ++jj)
        for (int ii = 0; ii < 2; ++ii) {
            double wxy = wx[ii] * wy[jj];
            size_t k = (ix[ii] - 1) + (size_t)(iy[jj] - 1) * NX;
            a = fma(Ex[k], wxy, a);
            b = fma(Ey[k], wxy, b);
        }
    ex[p] = a; ey[p] = b;
}

__global__ void stage_boris_kernel(double *vx, double *vy, double *vz, const double *Ex, const double *Ey, long long count,
                                   double dt, double t1, double tscale)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    double a = vx[p], b = vy[p], c = vz[p];
    boris(a, b, c, Ex[p], Ey[p], dt, t1, tscale);
    vx[p] = a; vy[p] = b; vz[p] = c;
}

__global__ void split_E2_kernel(const double2 *E2, long long n, double *Ex, double *Ey)
{
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { double2 e = E2[k]; if (Ex) Ex[k] = e.x; if (Ey) Ey[k] = e.y; }
}

} // namespace pg
