// Test infrastructure (NOT part of the product): compiles particleincellcodegolf.jl_b200/csrc/pg_es_math.h -- the very
// header the CUDA kernels of pg_esfield.cuh evaluate on the device -- for the host with g++ -ffp-contract=off, so that
// tests/test_esfield_cpu.py can compare the arithmetic with the oracle bit for bit without a GPU, and a serial
// emulation of es_particles_kernel (periodic grid, unimod-wrapped stencil indices, 64-bit fixed-point charge) can be
// checked against the oracle's literal halo-array restatement of PIC2D3V.jl:530-560.
#include <stdint.h>
#include <math.h>
#include "pg_es_math.h"

using namespace pg::es;

extern "C" {

int h_shape(int shape, double z, double NZ_Lz, int *j0, double *w6) { return shape_weights_rt(shape, z, NZ_Lz, *j0, w6); }

void h_boris(double *v, double Ex, double Ey, const double *B, double dt, double q_m)
{
    Boris b = make_boris(B[0], B[1], B[2], dt, q_m);
    boris_push(b, v[0], v[1], v[2], Ex, Ey);
}

double h_halton(long long i, int base, double seed) { return halton(i, base, seed); }

// The loop body of es_particles_kernel for one species, serially.  Exy2: interleaved (x, y) pairs per periodic cell.
// rho_fx accumulates round(wx*wy*dep * fx_scale) like the kernel's integer REDs.  sums[7] as the kernel's partials.
void h_emulate_particles(int shape, long long P, double *x, double *y, double *vx, double *vy, double *vz, const double *Exy2,
                         long long *rho_fx, int NX, int NY, double Lx, double Ly, double dt, const double *B, double q_m, double dep,
                         double fx_scale, double *sums)
{
    const Boris boris = make_boris(B[0], B[1], B[2], dt, q_m);
    const double NX_Lx = (double)NX / Lx, NY_Ly = (double)NY / Ly;
    const int S = support(shape);
    for (int k = 0; k < 7; ++k) sums[k] = 0.0;
    for (long long p = 0; p < P; ++p) {
        int ix0, iy0, cx[6], cy[6];
        double wx[6], wy[6];
        shape_weights_rt(shape, x[p], NX_Lx, ix0, wx);
        shape_weights_rt(shape, y[p], NY_Ly, iy0, wy);
        for (int s = 0; s < S; ++s) { cx[s] = unimod(ix0 + s, NX) - 1; cy[s] = (unimod(iy0 + s, NY) - 1) * NX; }
        double Exi = 0.0, Eyi = 0.0;
        for (int jj = 0; jj < S; ++jj)
            for (int ii = 0; ii < S; ++ii) {
                const double wxy = wx[ii] * wy[jj];
                Exi = fma(Exy2[2 * (cx[ii] + cy[jj])], wxy, Exi);
                Eyi = fma(Exy2[2 * (cx[ii] + cy[jj]) + 1], wxy, Eyi);
            }
        const double vxi = vx[p], vyi = vy[p];
        boris_push(boris, vx[p], vy[p], vz[p], Exi, Eyi);
        x[p] = unimod(x[p] + (vxi + vx[p]) / 2 * dt, Lx);
        y[p] = unimod(y[p] + (vyi + vy[p]) / 2 * dt, Ly);
        shape_weights_rt(shape, x[p], NX_Lx, ix0, wx);
        shape_weights_rt(shape, y[p], NY_Ly, iy0, wy);
        for (int s = 0; s < S; ++s) { cx[s] = unimod(ix0 + s, NX) - 1; cy[s] = (unimod(iy0 + s, NY) - 1) * NX; }
        for (int jj = 0; jj < S; ++jj)
            for (int ii = 0; ii < S; ++ii) rho_fx[cx[ii] + cy[jj]] += llrint(wx[ii] * wy[jj] * dep * fx_scale);
        sums[0] += vx[p] * vx[p] + vy[p] * vy[p] + vz[p] * vz[p];
        sums[1] += vx[p]; sums[2] += vy[p]; sums[3] += vz[p];
        sums[4] += fabs(vx[p]); sums[5] += fabs(vy[p]); sums[6] += fabs(vz[p]);
    }
}

} // extern "C"
