// pg_esfield.cuh -- kernels of the PIC2D3V.jl electrostatic path (include/picgolf_es.h):
//   loop!(plasma, field::ElectrostaticField, ...)      src/PIC2D3V.jl:530-581
//   update!(f::ElectrostaticField)                      src/PIC2D3V.jl:294-297
//   diagnose!(d::ElectrostaticDiagnostics, f, plasma)   src/PIC2D3V.jl:1301-1330
// Generalises the Electrostatic2D3V.jl loop (pg_kernels_2d.cuh) to several species (charge, mass, weight), shapes
// NGP / area / B-spline 0..5, Lx, Ly != 1 and a magnetic field with three components.
//
// The reference gives every thread a halo ("offset") copy of the charge grid (buffer 3), deposits with unwrapped
// indices and folds the halos back with applyperiodicity! (:13-19); its gather reads a halo copy of the field that
// update! fills from the periodic grid (:21-27).  Here one periodic grid per quantity lives in HBM and stencil indices
// are wrapped with the same unimod -- the same sums, term by term.  Charge accumulates in 64-bit fixed point (integer
// REDs: order-free and bit-reproducible for any particle order and GPU count, as in the other schemes).
#pragma once
#include "pg_common.cuh"
#include "pg_es_math.h"
#include "pg_fft.cuh"
#include "pg_kernels_2d.cuh"
#include "pg_sort.cuh"

namespace pg {

constexpr int ES_NSUM = 7;      // per-species particle sums: sum(abs2, v), sum(v)[3], sum(abs.(v))[3]
constexpr int ES_NSCALAR = 8;   // one diagnostics row: kinetic, field, particlemomentum[3], characteristicmomentum[3]
constexpr int ES_MAXSPECIES = 4;
constexpr int ES_BUFFER = 3;    // halo width of ElectrostaticField (:284): enters the field energy, see es_solve_rows_inv

struct ESParticleArgs {
    double *x, *y, *vx, *vy, *vz;   // this species' local shard
    long long P;
    const double2 *Exy;             // (Exy[1,i,j], Exy[2,i,j]) per cell, column-major NX x NY (periodic image of the halo array)
    fx_t *rho;                      // fixed-point deposit grid: sum of wx*wy*dep
    double *partials;               // [ES_NSUM * gridDim.x]
    es::Boris boris;
    double NX_Lx, NY_Ly, Lx, Ly, dt, q_m;
    double dep;                     // qw_dV / wref (|dep| <= 1): the solve multiplies by wref
    double fx_scale;
    int NX, NY;
    // tile-sorted mode (es_particles_tiled): as P2DArgs
    const unsigned int *tile_start, *tile_end; // particle range of each tile in the sorted arrays
    const unsigned int *item_off;              // [ntiles+1] prefix of ceil(count/T2_CHUNK): work items per tile
    unsigned long long *slow_count;
    double fxw_scale;    // 2^fracw: format of the shared-memory window, which accumulates the UNSIGNED fractions wx*wy
    double wscale;       // window -> global: dep * 2^(frac - fracw), applied once per window cell at the flush
    int ntx, ntiles;
};

// 0-based periodic cell of the (1-based, unwrapped) stencil index i: the reference's single-wrap unimod; an index that
// is still outside 1..N after it (a particle that left the box by more than a period -- the reference would throw a
// BoundsError on its halo array -- or a NaN position) is folded with a full modulo instead of touching foreign memory.
__device__ __forceinline__ int es_cell0(int i, int N)
{
    int u = es::unimod(i, N) - 1;
    if ((unsigned)u >= (unsigned)N) { u = (i - 1) % N; if (u < 0) u += N; }
    return u;
}

// for i in species.chunks[k] (:546-553): gather -> boris -> move -> deposit, any particle order.
template <int SHAPE>
__global__ void __launch_bounds__(PG_THREADS) es_particles_kernel(ESParticleArgs a)
{
    constexpr int S = es::support(SHAPE);
    __shared__ double scratch[32];
    const int NX = a.NX, NY = a.NY;
    double sum[ES_NSUM];
#pragma unroll
    for (int k = 0; k < ES_NSUM; ++k) sum[k] = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.P; p += stride) {
        double x = ld_stream(a.x + p), y = ld_stream(a.y + p);
        double vx = ld_stream(a.vx + p), vy = ld_stream(a.vy + p), vz = ld_stream(a.vz + p);
        int ix0, iy0, cx[S], cy[S];
        double wx[S], wy[S];
        es::shape_weights<SHAPE>(x, a.NX_Lx, ix0, wx);
        es::shape_weights<SHAPE>(y, a.NY_Ly, iy0, wy);
#pragma unroll
        for (int s = 0; s < S; ++s) { cx[s] = es_cell0(ix0 + s, NX); cy[s] = es_cell0(iy0 + s, NY) * NX; }
        double Exi = 0.0, Eyi = 0.0; // field(species.shape, x[i], y[i])   :1217-1229 (j outer, i inner, @muladd)
#pragma unroll
        for (int jj = 0; jj < S; ++jj)
#pragma unroll
            for (int ii = 0; ii < S; ++ii) {
                const double wxy = wx[ii] * wy[jj];
                const double2 f = __ldg(&a.Exy[cx[ii] + cy[jj]]);
                Exi = fma(f.x, wxy, Exi);
                Eyi = fma(f.y, wxy, Eyi);
            }
        const double vxi = vx, vyi = vy;
        es::boris_push(a.boris, vx, vy, vz, Exi, Eyi);
        x = es::unimod(x + (vxi + vx) / 2 * a.dt, a.Lx);
        y = es::unimod(y + (vyi + vy) / 2 * a.dt, a.Ly);
        es::shape_weights<SHAPE>(x, a.NX_Lx, ix0, wx);
        es::shape_weights<SHAPE>(y, a.NY_Ly, iy0, wy);
#pragma unroll
        for (int s = 0; s < S; ++s) { cx[s] = es_cell0(ix0 + s, NX); cy[s] = es_cell0(iy0 + s, NY) * NX; }
#pragma unroll
        for (int jj = 0; jj < S; ++jj)
#pragma unroll
            for (int ii = 0; ii < S; ++ii) // deposit!: z[i,j] += wx * wy * w   :1246-1253
                atomicAdd(&a.rho[cx[ii] + cy[jj]], to_fx(wx[ii] * wy[jj] * a.dep, a.fx_scale));
        st_stream(a.x + p, x); st_stream(a.y + p, y);
        st_stream(a.vx + p, vx); st_stream(a.vy + p, vy); st_stream(a.vz + p, vz);
        sum[0] += vx * vx + vy * vy + vz * vz;
        sum[1] += vx; sum[2] += vy; sum[3] += vz;
        sum[4] += fabs(vx); sum[5] += fabs(vy); sum[6] += fabs(vz);
    }
#pragma unroll
    for (int k = 0; k < ES_NSUM; ++k) {
        const double s = block_sum(sum[k], scratch);
        if (threadIdx.x == 0) a.partials[ES_NSUM * blockIdx.x + k] = s;
    }
}

// Tile-sorted variant (the scheme of particles_2d3v_tiled, pg_kernels_2d.cuh, for any shape).  The particles of a species
// are kept sorted by T2_TS x T2_TS-cell tile (pg_sort.cuh, mode 2); a work item is up to T2_CHUNK consecutive particles of
// ONE tile.  The block stages the T2_WS x T2_WS window of Exy around the tile in shared memory and accumulates the
// fractions wx*wy of the deposit into a fixed-point window of the same size -- two 32-bit limbs with native shared-memory
// atomics (64-bit ones are CAS loops on sm_100a); the species' signed weight q*w/dV enters once per window cell when the
// window is flushed with integer REDs.  So the S^2 gathers and S^2 deposits of a particle never leave the SM.  A stencil
// that does not fit the window (a particle that drifted more than T2_R - S cells out of its tile since the last sort) goes
// to global memory and is counted.
#ifndef PG_ES_MINBLOCKS
#define PG_ES_MINBLOCKS 2 // resident blocks per SM the register allocation aims at (A/B builds: tools/build_variants.sh)
#endif
#ifndef PG_ES_PREFETCH
#define PG_ES_PREFETCH 1  // 1: the next particle's five values are loaded before the current one is processed
#endif
template <int SHAPE>
__global__ void __launch_bounds__(PG_THREADS, PG_ES_MINBLOCKS) es_particles_tiled(ESParticleArgs a)
{
    constexpr int S = es::support(SHAPE);
    __shared__ double2 Ew[T2_WS * T2_WS];
    __shared__ unsigned int rlo[T2_WS * T2_WS], rhi[T2_WS * T2_WS];
    __shared__ double scratch[32];
    const int NX = a.NX, NY = a.NY, mx = NX - 1, my = NY - 1;
    double sum[ES_NSUM];
#pragma unroll
    for (int k = 0; k < ES_NSUM; ++k) sum[k] = 0.0;
    unsigned int nslow = 0;
    const unsigned int nitems = a.item_off[a.ntiles];
    for (unsigned int item = blockIdx.x; item < nitems; item += gridDim.x) {
        int lo = 0, hi = a.ntiles; // tile of this item: last t with item_off[t] <= item
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (a.item_off[mid] <= item) lo = mid; else hi = mid;
        }
        const int tile = lo;
        const long long start = (long long)a.tile_start[tile] + (long long)(item - a.item_off[tile]) * T2_CHUNK;
        const long long end = min(start + (long long)T2_CHUNK, (long long)a.tile_end[tile]);
        const int ox = (tile % a.ntx) * T2_TS - T2_R, oy = (tile / a.ntx) * T2_TS - T2_R; // window origin (0-based cells)
        for (int c = threadIdx.x; c < T2_WS * T2_WS; c += blockDim.x) {
            int gx = (ox + (c & (T2_WS - 1))) & mx, gy = (oy + (c >> 5)) & my;
            Ew[c] = a.Exy[gx + (size_t)gy * NX];
            rlo[c] = 0u; rhi[c] = 0u;
        }
        __syncthreads();
#if PG_ES_PREFETCH
        double nx = 0.0, ny = 0.0, nvx = 0.0, nvy = 0.0, nvz = 0.0;
        if (start + threadIdx.x < end) {
            const long long p0 = start + threadIdx.x;
            nx = ld_stream(a.x + p0); ny = ld_stream(a.y + p0); nvx = ld_stream(a.vx + p0); nvy = ld_stream(a.vy + p0); nvz = ld_stream(a.vz + p0);
        }
#endif
        for (long long p = start + threadIdx.x; p < end; p += blockDim.x) {
#if PG_ES_PREFETCH
            double x = nx, y = ny, vx = nvx, vy = nvy, vz = nvz;
            {
                const long long pn = p + blockDim.x;
                if (pn < end) { nx = ld_stream(a.x + pn); ny = ld_stream(a.y + pn); nvx = ld_stream(a.vx + pn); nvy = ld_stream(a.vy + pn); nvz = ld_stream(a.vz + pn); }
            }
#else
            double x = ld_stream(a.x + p), y = ld_stream(a.y + p);
            double vx = ld_stream(a.vx + p), vy = ld_stream(a.vy + p), vz = ld_stream(a.vz + p);
#endif
            int ix0, iy0;
            double wx[S], wy[S];
            es::shape_weights<SHAPE>(x, a.NX_Lx, ix0, wx);
            es::shape_weights<SHAPE>(y, a.NY_Ly, iy0, wy);
            double Exi = 0.0, Eyi = 0.0;
            {
                const int rx = (ix0 - 1 - ox) & mx, ry = (iy0 - 1 - oy) & my;
                if (rx <= T2_WS - S && ry <= T2_WS - S) {
                    const double2 *e = Ew + rx + ry * T2_WS;
#pragma unroll
                    for (int jj = 0; jj < S; ++jj)
#pragma unroll
                        for (int ii = 0; ii < S; ++ii) {
                            const double wxy = wx[ii] * wy[jj];
                            const double2 f = e[ii + jj * T2_WS];
                            Exi = fma(f.x, wxy, Exi);
                            Eyi = fma(f.y, wxy, Eyi);
                        }
                } else {
#pragma unroll
                    for (int jj = 0; jj < S; ++jj)
#pragma unroll
                        for (int ii = 0; ii < S; ++ii) {
                            const double wxy = wx[ii] * wy[jj];
                            const double2 f = __ldg(&a.Exy[es_cell0(ix0 + ii, NX) + (size_t)es_cell0(iy0 + jj, NY) * NX]);
                            Exi = fma(f.x, wxy, Exi);
                            Eyi = fma(f.y, wxy, Eyi);
                        }
                }
            }
            const double vxi = vx, vyi = vy;
            es::boris_push(a.boris, vx, vy, vz, Exi, Eyi);
            x = es::unimod(x + (vxi + vx) / 2 * a.dt, a.Lx);
            y = es::unimod(y + (vyi + vy) / 2 * a.dt, a.Ly);
            es::shape_weights<SHAPE>(x, a.NX_Lx, ix0, wx);
            es::shape_weights<SHAPE>(y, a.NY_Ly, iy0, wy);
            {
                const int rx = (ix0 - 1 - ox) & mx, ry = (iy0 - 1 - oy) & my;
                if (rx <= T2_WS - S && ry <= T2_WS - S) {
                    const int r0 = rx + ry * T2_WS;
#pragma unroll
                    for (int jj = 0; jj < S; ++jj)
#pragma unroll
                        for (int ii = 0; ii < S; ++ii) {
                            // the fractions of the B-splines of order >= 4 can come out as -1e-17 instead of 0 (cancellation in
                            // the reference's own polynomials): the two-limb add below is exact for either sign (mod 2^64)
                            const fx_t v = to_fx(wx[ii] * wy[jj], a.fxw_scale);
                            const unsigned int vlo = (unsigned int)v;
                            const unsigned int old = atomicAdd(&rlo[r0 + ii + jj * T2_WS], vlo);
                            const unsigned int carry = (old + vlo) < old ? 1u : 0u;
                            const unsigned int vhi = (unsigned int)(v >> 32) + carry;
                            if (vhi) atomicAdd(&rhi[r0 + ii + jj * T2_WS], vhi);
                        }
                } else {
                    ++nslow;
#pragma unroll
                    for (int jj = 0; jj < S; ++jj)
#pragma unroll
                        for (int ii = 0; ii < S; ++ii)
                            atomicAdd(&a.rho[es_cell0(ix0 + ii, NX) + (size_t)es_cell0(iy0 + jj, NY) * NX],
                                      to_fx(wx[ii] * wy[jj] * a.dep, a.fx_scale));
                }
            }
            st_stream(a.x + p, x); st_stream(a.y + p, y);
            st_stream(a.vx + p, vx); st_stream(a.vy + p, vy); st_stream(a.vz + p, vz);
            sum[0] += vx * vx + vy * vy + vz * vz;
            sum[1] += vx; sum[2] += vy; sum[3] += vz;
            sum[4] += fabs(vx); sum[5] += fabs(vy); sum[6] += fabs(vz);
        }
        __syncthreads();
        for (int c = threadIdx.x; c < T2_WS * T2_WS; c += blockDim.x) {
            const long long v = (long long)(((fx_t)rhi[c] << 32) | (fx_t)rlo[c]);
            if (v) {
                int gx = (ox + (c & (T2_WS - 1))) & mx, gy = (oy + (c >> 5)) & my;
                atomicAdd(&a.rho[gx + (size_t)gy * NX], (fx_t)__double2ll_rn((double)v * a.wscale));
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < ES_NSUM; ++k) {
        const double s = block_sum(sum[k], scratch);
        if (threadIdx.x == 0) a.partials[ES_NSUM * blockIdx.x + k] = s;
    }
    if (nslow && a.slow_count) atomicAdd(a.slow_count, (unsigned long long)nslow);
}

// ---------------------------------------------------------------------------------------------
// Slice-streaming variant (the scheme of particles_2d3v_stream, pg_kernels_2d.cuh, for any shape; default of the tile-sorted
// path).  One block of 512 threads per SM owns a contiguous slice of a species' tile-sorted arrays and walks through it tile
// segment by tile segment; rows of 64 particles (two per lane) arrive through per-warp 16-byte cp.async rings and leave with
// 128-bit stores; the field and deposit windows are replicated (lane l gathers from replica l % G and adds into replica
// l % D), which removes most of the bank conflicts of the S^2 gathers and 2 S^2 limb adds per particle.  The two particles
// of a lane are pushed together (independent gathers overlap) for supports up to 3; wider stencils run one after the other.
// ---------------------------------------------------------------------------------------------
template <int S>
struct ESStencil { int ix0, iy0; double wx[S], wy[S]; };

struct ESWindow {
    const double2 *Ew;
    unsigned int *rlo, *rhi;
    int ox, oy;
};

template <int SHAPE, int G>
__device__ __forceinline__ void es_push(const ESParticleArgs &a, const ESWindow &w, double &x, double &y, double &vx, double &vy,
                                        double &vz, double *sum, ESStencil<es::support(SHAPE)> &st)
{
    constexpr int S = es::support(SHAPE);
    const int NX = a.NX, NY = a.NY, mx = NX - 1, my = NY - 1;
    es::shape_weights<SHAPE>(x, a.NX_Lx, st.ix0, st.wx);
    es::shape_weights<SHAPE>(y, a.NY_Ly, st.iy0, st.wy);
    double Exi = 0.0, Eyi = 0.0;
    {
        const int rx = (st.ix0 - 1 - w.ox) & mx, ry = (st.iy0 - 1 - w.oy) & my;
        if (rx <= T2_WS - S && ry <= T2_WS - S) {
            const double2 *e = w.Ew + (rx + ry * T2_WS) * G;
#pragma unroll
            for (int jj = 0; jj < S; ++jj)
#pragma unroll
                for (int ii = 0; ii < S; ++ii) {
                    const double wxy = st.wx[ii] * st.wy[jj];
                    const double2 f = e[(ii + jj * T2_WS) * G];
                    Exi = fma(f.x, wxy, Exi);
                    Eyi = fma(f.y, wxy, Eyi);
                }
        } else {
#pragma unroll
            for (int jj = 0; jj < S; ++jj)
#pragma unroll
                for (int ii = 0; ii < S; ++ii) {
                    const double wxy = st.wx[ii] * st.wy[jj];
                    const double2 f = __ldg(&a.Exy[es_cell0(st.ix0 + ii, NX) + (size_t)es_cell0(st.iy0 + jj, NY) * NX]);
                    Exi = fma(f.x, wxy, Exi);
                    Eyi = fma(f.y, wxy, Eyi);
                }
        }
    }
    const double vxi = vx, vyi = vy;
    es::boris_push(a.boris, vx, vy, vz, Exi, Eyi);
    x = es::unimod(x + (vxi + vx) / 2 * a.dt, a.Lx);
    y = es::unimod(y + (vyi + vy) / 2 * a.dt, a.Ly);
    es::shape_weights<SHAPE>(x, a.NX_Lx, st.ix0, st.wx);
    es::shape_weights<SHAPE>(y, a.NY_Ly, st.iy0, st.wy);
    sum[0] += vx * vx + vy * vy + vz * vz;
    sum[1] += vx; sum[2] += vy; sum[3] += vz;
    sum[4] += fabs(vx); sum[5] += fabs(vy); sum[6] += fabs(vz);
}

template <int S, int D>
__device__ __forceinline__ void es_deposit(const ESParticleArgs &a, const ESWindow &w, const ESStencil<S> &st, unsigned int &nslow)
{
    const int NX = a.NX, NY = a.NY, mx = NX - 1, my = NY - 1;
    const int rx = (st.ix0 - 1 - w.ox) & mx, ry = (st.iy0 - 1 - w.oy) & my;
    if (rx <= T2_WS - S && ry <= T2_WS - S) {
        const int r0 = (rx + ry * T2_WS) * D;
#pragma unroll
        for (int jj = 0; jj < S; ++jj)
#pragma unroll
            for (int ii = 0; ii < S; ++ii) { // two-limb add, exact for either sign (mod 2^64); see es_particles_tiled
                const fx_t v = to_fx_small(st.wx[ii] * st.wy[jj], a.fxw_scale); // fractions <= 1, fxw_scale <= 2^48
                const unsigned int vlo = (unsigned int)v;
                const unsigned int old = atomicAdd(&w.rlo[r0 + (ii + jj * T2_WS) * D], vlo);
                const unsigned int carry = (old + vlo) < old ? 1u : 0u;
                const unsigned int vhi = (unsigned int)(v >> 32) + carry;
                if (vhi) atomicAdd(&w.rhi[r0 + (ii + jj * T2_WS) * D], vhi);
            }
    } else {
        ++nslow;
#pragma unroll
        for (int jj = 0; jj < S; ++jj)
#pragma unroll
            for (int ii = 0; ii < S; ++ii)
                atomicAdd(&a.rho[es_cell0(st.ix0 + ii, NX) + (size_t)es_cell0(st.iy0 + jj, NY) * NX],
                          to_fx(st.wx[ii] * st.wy[jj] * a.dep, a.fx_scale));
    }
}

template <int SHAPE, int G, int D, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) es_particles_stream(ESParticleArgs a)
{
    constexpr int S = es::support(SHAPE), NC = T2_WS * T2_WS, nw = THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *Ew = reinterpret_cast<double2 *>(smem_raw);                  // [cell][G]
    unsigned int *rlo = reinterpret_cast<unsigned int *>(Ew + NC * G), *rhi = rlo + NC * D; // [cell][D]
    double *scratch = reinterpret_cast<double *>(rhi + NC * D);
    double2 *ring = reinterpret_cast<double2 *>(scratch + 32);            // [warp][stage][array][lane]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double2 *const mine = ring + (size_t)wid * S2_STAGES * S2_ROW + lane;
    const int NX = a.NX, NY = a.NY, mx = NX - 1, my = NY - 1;
    double *const gp[5] = {a.x, a.y, a.vx, a.vy, a.vz};
    double sum[ES_NSUM];
#pragma unroll
    for (int k = 0; k < ES_NSUM; ++k) sum[k] = 0.0;
    unsigned int nslow = 0;
    ESWindow w;
    w.Ew = Ew + (lane & (G - 1)); w.rlo = rlo + (lane & (D - 1)); w.rhi = rhi + (lane & (D - 1)); w.ox = 0; w.oy = 0;
    for (int c = threadIdx.x; c < NC * D; c += THREADS) { rlo[c] = 0u; rhi[c] = 0u; }
    const long long per = (((a.P + gridDim.x - 1) / gridDim.x) + 63) & ~63LL;
    const long long s_lo = min(a.P, per * (long long)blockIdx.x), s_hi = min(a.P, s_lo + per);
    int tile = 0;
    {
        int lo = 0, hi = a.ntiles; // last t with tile_start[t] <= s_lo
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if ((long long)a.tile_start[mid] <= s_lo) lo = mid; else hi = mid;
        }
        tile = lo;
    }
    int cur_tile = -1;
    for (long long pos = s_lo; pos < s_hi;) {
        while (tile < a.ntiles - 1 && (long long)a.tile_end[tile] <= pos) ++tile;
        const long long seg_hi = min(min(s_hi, (long long)a.tile_end[tile]), pos + S2_FLUSH);
        const long long base = pos & ~63LL;
        const int nrows = (int)((seg_hi - base + 63) >> 6);
        auto issue = [&](int k) { // lane copies the pair (p0, p0+1) of each array; always commits (uniform group count)
            const int r = wid + k * nw;
            const long long p0 = base + ((long long)r << 6) + 2 * lane;
            if (r < nrows && p0 + 1 >= pos && p0 < seg_hi) {
                double2 *st = mine + (k % S2_STAGES) * S2_ROW;
                if (p0 + 1 < a.P) {
#pragma unroll
                    for (int q = 0; q < 5; ++q) cp_async16(st + 32 * q, gp[q] + p0);
                } else { // the very last particle of an odd-sized shard
#pragma unroll
                    for (int q = 0; q < 5; ++q) cp_async8(st + 32 * q, gp[q] + p0);
                }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int k = 0; k < S2_STAGES - 1; ++k) issue(k);
        if (tile != cur_tile) { // (re)build the field window around this tile, all replicas
            w.ox = (tile % a.ntx) * T2_TS - T2_R; w.oy = (tile / a.ntx) * T2_TS - T2_R;
            for (int c = threadIdx.x; c < NC * G; c += THREADS) {
                const int cell = c / G;
                int gx = (w.ox + (cell & (T2_WS - 1))) & mx, gy = (w.oy + (cell >> 5)) & my;
                Ew[c] = a.Exy[gx + (size_t)gy * NX];
            }
            cur_tile = tile;
        }
        __syncthreads();
        for (int k = 0; wid + k * nw < nrows; ++k) {
            issue(k + S2_STAGES - 1);
            cp_async_wait<S2_STAGES - 1>(); // row k has landed
            const long long p0 = base + ((long long)(wid + k * nw) << 6) + 2 * lane;
            const bool v0 = p0 >= pos && p0 < seg_hi, v1 = p0 + 1 >= pos && p0 + 1 < seg_hi;
            if (v0 || v1) {
                const double2 *st = mine + (k % S2_STAGES) * S2_ROW;
                double2 X = st[0], Y = st[32], VX = st[64], VY = st[96], VZ = st[128];
                if (S <= 3 && v0 && v1) { // the two pushes are independent: their gathers overlap
                    ESStencil<S> c0, c1;
                    es_push<SHAPE, G>(a, w, X.x, Y.x, VX.x, VY.x, VZ.x, sum, c0);
                    es_push<SHAPE, G>(a, w, X.y, Y.y, VX.y, VY.y, VZ.y, sum, c1);
                    es_deposit<S, D>(a, w, c0, nslow);
                    es_deposit<S, D>(a, w, c1, nslow);
                } else {
                    ESStencil<S> c;
                    if (v0) { es_push<SHAPE, G>(a, w, X.x, Y.x, VX.x, VY.x, VZ.x, sum, c); es_deposit<S, D>(a, w, c, nslow); }
                    if (v1) { es_push<SHAPE, G>(a, w, X.y, Y.y, VX.y, VY.y, VZ.y, sum, c); es_deposit<S, D>(a, w, c, nslow); }
                }
                if (v0 && v1) {
                    __stcs(reinterpret_cast<double2 *>(a.x + p0), X); __stcs(reinterpret_cast<double2 *>(a.y + p0), Y);
                    __stcs(reinterpret_cast<double2 *>(a.vx + p0), VX); __stcs(reinterpret_cast<double2 *>(a.vy + p0), VY);
                    __stcs(reinterpret_cast<double2 *>(a.vz + p0), VZ);
                } else if (v0) {
                    st_stream(a.x + p0, X.x); st_stream(a.y + p0, Y.x);
                    st_stream(a.vx + p0, VX.x); st_stream(a.vy + p0, VY.x); st_stream(a.vz + p0, VZ.x);
                } else {
                    st_stream(a.x + p0 + 1, X.y); st_stream(a.y + p0 + 1, Y.y);
                    st_stream(a.vx + p0 + 1, VX.y); st_stream(a.vy + p0 + 1, VY.y); st_stream(a.vz + p0 + 1, VZ.y);
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        for (int c = threadIdx.x; c < NC; c += THREADS) { // flush: sum the D replicas of each window cell, clear them
            unsigned long long v = 0ULL;
#pragma unroll
            for (int r = 0; r < D; ++r) {
                v += ((fx_t)rhi[c * D + r] << 32) + (fx_t)rlo[c * D + r];
                rlo[c * D + r] = 0u; rhi[c * D + r] = 0u;
            }
            if (v) { // the species' signed weight enters here, once per window cell
                int gx = (w.ox + (c & (T2_WS - 1))) & mx, gy = (w.oy + (c >> 5)) & my;
                atomicAdd(&a.rho[gx + (size_t)gy * NX], (fx_t)__double2ll_rn((double)(long long)v * a.wscale));
            }
        }
        pos = seg_hi; // the __syncthreads() at the top of the next segment orders the clears before its deposits
    }
#pragma unroll
    for (int k = 0; k < ES_NSUM; ++k) {
        const double s = block_sum(sum[k], scratch);
        if (threadIdx.x == 0) a.partials[ES_NSUM * blockIdx.x + k] = s;
    }
    if (nslow && a.slow_count) atomicAdd(a.slow_count, (unsigned long long)nslow);
}

// ---------------------------------------------------------------------------------------------
// Field solve (:562-579): pfft! * phi; phi[1,1] = 0; tmp = phi * im_k^-2; Ex = tmp*kx[i]; Ey = tmp*ky[j]; pifft! both,
// with FFTHelper's kx = 2pi/Lx * vcat(0:NX/2-1, -NX/2:-1) (:254-258).  Pass A is solve2d_rows_fwd (pg_fft.cuh); these
// are pass B and C of the same three-kernel transform with the box lengths, and with what update!/diagnose! need.
// ---------------------------------------------------------------------------------------------
struct ESSolveArgs {
    double2 *Z;               // [NX*NY] spectrum scratch (x bit-reversed after pass A)
    double2 *Enow;            // (real(Ex), real(Ey)) of THIS solve: what diagnose! averages into Exs/Eys (:1323-1324)
    double2 *Exy;             // the field the particles gather from: update! ADDS Enow to it (:294-297), see `accumulate`
    const double *rho_last;   // charge density of this solve (pass A wrote it)
    const double2 *twx, *twy;
    double *epartials;        // per-block partials of sum(abs2, Exy) over the halo array
    double *dc;               // [1] sum(rho) (the [1,1] entry of the spectrum before it is zeroed)
    double *hist_ex, *hist_ey, *hist_phi; // slice ti of Exs/Eys/phis (NXd x NYd, column-major) or nullptr
    double kxs, kys;          // 2pi/Lx, 2pi/Ly
    double ntskip;            // d.ntskip as a double
    int NX, NY, lgx, lgy, accumulate, ngskip;
};

__global__ void __launch_bounds__(512) es_solve_cols(ESSolveArgs a)
{
    extern __shared__ double smem[];
    const int NX = a.NX, NY = a.NY, C = COLS_PER_BLOCK, LD = NY + 1;
    double *re = smem, *im = smem + C * LD;
    const int p0 = blockIdx.x * C;
    for (int t = threadIdx.x; t < C * NY; t += blockDim.x) {
        int j = t / C, c = t - j * C;
        double2 z = a.Z[(size_t)(p0 + c) + (size_t)j * NX];
        re[c * LD + j] = z.x; im[c * LD + j] = z.y;
    }
    __syncthreads();
    fft_smem4<false>(re, im, NY, C, 1, LD, a.twy, NY);
    for (int t = threadIdx.x; t < C * NY; t += blockDim.x) {
        int c = t / NY, q = t - c * NY;
        int ix = bitrev(p0 + c, a.lgx), iy = bitrev(q, a.lgy); // frequency slots (0-based)
        double kx = a.kxs * (double)(ix < NX / 2 ? ix : ix - NX);
        double ky = a.kys * (double)(iy < NY / 2 ? iy : iy - NY);
        double ar = re[c * LD + q], ai = im[c * LD + q];
        double zr = 0.0, zi = 0.0;
        if (ix != 0 || iy != 0) {
            double m = -1.0 / (kx * kx + ky * ky); // im_k^-2 = (0, m)
            double tr = -(ai * m), ti = ar * m;    // phi * im_k^-2
            double exr = tr * kx, exi = ti * kx, eyr = tr * ky, eyi = ti * ky;
            // update! keeps real(ifft(.)) only: the non-Hermitian Nyquist row/column drops out (see pg_fft.cuh)
            if (ix == NX / 2) { exr = 0.0; exi = 0.0; }
            if (iy == NY / 2) { eyr = 0.0; eyi = 0.0; }
            zr = exr - eyi; zi = exi + eyr; // Ex^ + i Ey^
        } else {
            *a.dc = ar;
        }
        re[c * LD + q] = zr; im[c * LD + q] = zi;
    }
    __syncthreads();
    fft_smem4<true>(re, im, NY, C, 1, LD, a.twy, NY);
    for (int t = threadIdx.x; t < C * NY; t += blockDim.x) {
        int j = t / C, c = t - j * C;
        a.Z[(size_t)(p0 + c) + (size_t)j * NX] = make_double2(re[c * LD + j], im[c * LD + j]);
    }
}

// How many entries of the halo array (indices -(B-1) : NZ+B, B = ES_BUFFER) map to periodic cell i (1-based) under unimod.
__device__ __forceinline__ int es_halo_multiplicity(int i, int NZ) { return 1 + (i <= ES_BUFFER ? 1 : 0) + (i > NZ - ES_BUFFER ? 1 : 0); }

__global__ void __launch_bounds__(512) es_solve_rows_inv(ESSolveArgs a)
{
    extern __shared__ double smem[];
    const int NX = a.NX, NY = a.NY, R = ROWS_PER_BLOCK;
    double *re = smem, *im = smem + R * NX, *scratch = smem + 2 * R * NX;
    const int j0 = blockIdx.x * R;
    for (int t = threadIdx.x; t < R * NX; t += blockDim.x) {
        int r = t / NX, i = t - r * NX;
        double2 z = a.Z[(size_t)i + (size_t)(j0 + r) * NX];
        re[t] = z.x; im[t] = z.y;
    }
    __syncthreads();
    fft_smem4<true>(re, im, NX, R, 1, NX, a.twx, NX);
    const double inv = (double)NX * (double)NY;
    const double mean = *a.dc / inv;
    const int g = a.ngskip, NXd = NX / g;
    double e2 = 0.0;
    for (int t = threadIdx.x; t < R * NX; t += blockDim.x) {
        int r = t / NX, i = t - r * NX, j = j0 + r;
        size_t cell = (size_t)i + (size_t)j * NX;
        double ex = re[t] / inv, ey = im[t] / inv;
        a.Enow[cell] = make_double2(ex, ey);
        double2 f = make_double2(ex, ey);
        if (a.accumulate) { double2 o = a.Exy[cell]; f.x = o.x + ex; f.y = o.y + ey; } // oa[i,j] += real(a[...])   :24-26
        a.Exy[cell] = f;
        // mean(abs2, f.Exy) runs over the halo array: a periodic cell within ES_BUFFER of an edge appears twice per dimension
        e2 += (double)(es_halo_multiplicity(i + 1, NX) * es_halo_multiplicity(j + 1, NY)) * (f.x * f.x + f.y * f.y);
        if (a.hist_ex && (i % g) == 0 && (j % g) == 0) { // d.Exs[:, :, ti] .+= real.(f.Ex[a, b]) ./ d.ntskip   :1323-1325
            size_t hc = (size_t)(i / g) + (size_t)(j / g) * NXd;
            a.hist_ex[hc] += ex / a.ntskip;
            a.hist_ey[hc] += ey / a.ntskip;
            // real.(pifft! * phi): phi is the spectrum of rho with [1,1] zeroed, i.e. rho - mean(rho)
            a.hist_phi[hc] += (a.rho_last[cell] - mean) / a.ntskip;
        }
    }
    e2 = block_sum(e2, scratch);
    if (threadIdx.x == 0) a.epartials[blockIdx.x] = e2;
}

// diagnose! (:1301-1321): one row of scalars from the per-species particle sums of this rank's shard and the field sum.
struct ESStepEndArgs {
    const double *partials;   // [nspecies][maxblocks][ES_NSUM]
    const double *epartials;  // [neblocks]
    double *rows;             // [ND][ES_NSCALAR]
    double mass[ES_MAXSPECIES], weight[ES_MAXSPECIES];
    int nblocks[ES_MAXSPECIES];
    int nspecies, maxblocks, neblocks, row; // row < 0: nothing to record this step
    double halo_count;        // 2 * (NX + 2B) * (NY + 2B): the length mean(abs2, Exy) divides by
};

__global__ void __launch_bounds__(256) es_step_end_kernel(ESStepEndArgs a)
{
    __shared__ double scratch[32];
    if (a.row < 0) return;
    double out[ES_NSCALAR];
#pragma unroll
    for (int k = 0; k < ES_NSCALAR; ++k) out[k] = 0.0;
    for (int s = 0; s < a.nspecies; ++s) {
        double sum[ES_NSUM];
#pragma unroll
        for (int k = 0; k < ES_NSUM; ++k) sum[k] = 0.0;
        for (int b = threadIdx.x; b < a.nblocks[s]; b += blockDim.x)
#pragma unroll
            for (int k = 0; k < ES_NSUM; ++k) sum[k] += a.partials[((size_t)s * a.maxblocks + b) * ES_NSUM + k];
#pragma unroll
        for (int k = 0; k < ES_NSUM; ++k) sum[k] = block_sum(sum[k], scratch);
        // kineticenergy(s) = sum(abs2, velocities(s)) * s.mass / 2 * s.weight   :177;  momentum(s, op) *= mass * weight   :185
        out[0] += sum[0] * a.mass[s] / 2 * a.weight[s];
#pragma unroll
        for (int k = 0; k < 3; ++k) { out[2 + k] += sum[1 + k] * (a.mass[s] * a.weight[s]); out[5 + k] += sum[4 + k] * (a.mass[s] * a.weight[s]); }
    }
    double e = 0.0;
    for (int b = threadIdx.x; b < a.neblocks; b += blockDim.x) e += a.epartials[b];
    e = block_sum(e, scratch);
    if (threadIdx.x == 0) {
        out[1] = e / a.halo_count / 2; // d.fieldenergy[ti] = mean(abs2, f.Exy) / 2   :1320
#pragma unroll
        for (int k = 0; k < ES_NSCALAR; ++k) a.rows[(size_t)a.row * ES_NSCALAR + k] = out[k];
    }
}

// ---------------------------------------------------------------------------------------------
// Species(P, vth, density, shape; Lx, Ly) (:194-213): Halton quiet start.  Positions and the raw velocities
// vth * erfinv(2 sample - 1) * vth are generated from the GLOBAL particle index (shards agree); the mean removal and the
// rescaling to std vth/sqrt(2) need global sums, done by es_moments_kernel / es_affine_kernel around an all-reduce.
// ---------------------------------------------------------------------------------------------
__global__ void es_species_init_kernel(double *x, double *y, double *vx, double *vy, double *vz, long long count, long long first,
                                       double Lx, double Ly, double vth)
{
    const double seed = 1 / sqrt(2.0);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < count; n += stride) {
        const long long i = first + n;
        x[n] = Lx * es::halton(i, 2, seed);
        y[n] = Ly * es::halton(i, 3, seed);
        vx[n] = vth * erfinv(2 * es::halton(i, 5, seed) - 1) * vth;
        vy[n] = vth * erfinv(2 * es::halton(i, 7, seed) - 1) * vth;
        vz[n] = vth * erfinv(2 * es::halton(i, 9, seed) - 1) * vth;
    }
}

// out[2*blockIdx.x + {0,1}] = sum(v - shift), sum((v - shift)^2) over this block's part of the shard
__global__ void __launch_bounds__(256) es_moments_kernel(const double *v, long long count, double shift, double *out)
{
    __shared__ double scratch[32];
    double s1 = 0.0, s2 = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < count; n += stride) {
        const double d = v[n] - shift;
        s1 += d; s2 += d * d;
    }
    s1 = block_sum(s1, scratch);
    s2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = s1; out[2 * blockIdx.x + 1] = s2; }
}

// out[0..1] = column sums of the nb x 2 block partials (one block)
__global__ void __launch_bounds__(256) es_reduce2_kernel(const double *partials, int nb, double *out)
{
    __shared__ double scratch[32];
    double s1 = 0.0, s2 = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) { s1 += partials[2 * b]; s2 += partials[2 * b + 1]; }
    s1 = block_sum(s1, scratch);
    s2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) { out[0] = s1; out[1] = s2; }
}

// v = (v - shift) * scale
__global__ void es_affine_kernel(double *v, long long count, double shift, double scale)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < count; n += stride) v[n] = (v[n] - shift) * scale;
}

// ---- stage kernels (parity tests call these through the C ABI) ------------------------------------------------
__global__ void es_stage_shape_kernel(int shape, const double *z, long long count, double NZ_Lz, int *j0, double *wt)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    double w[6] = {0, 0, 0, 0, 0, 0};
    int j = 0;
    es::shape_weights_rt(shape, z[p], NZ_Lz, j, w);
    j0[p] = j;
    for (int k = 0; k < 6; ++k) wt[6 * p + k] = w[k];
}

__global__ void es_stage_boris_kernel(double *vx, double *vy, double *vz, const double *Ex, const double *Ey, long long count,
                                      es::Boris b)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    double a = vx[p], c = vy[p], d = vz[p];
    es::boris_push(b, a, c, d, Ex[p], Ey[p]);
    vx[p] = a; vy[p] = c; vz[p] = d;
}

// xyv[5, P] (Julia column-major: particle-contiguous records of x, y, vx, vy, vz; Species.xyv :153) <-> five SoA streams
__global__ void es_xyv_to_soa_kernel(const double *xyv, long long count, double *x, double *y, double *vx, double *vy, double *vz)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < count; n += stride) {
        x[n] = xyv[5 * n]; y[n] = xyv[5 * n + 1]; vx[n] = xyv[5 * n + 2]; vy[n] = xyv[5 * n + 3]; vz[n] = xyv[5 * n + 4];
    }
}
__global__ void es_soa_to_xyv_kernel(const double *x, const double *y, const double *vx, const double *vy, const double *vz,
                                     long long count, double *xyv)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < count; n += stride) {
        xyv[5 * n] = x[n]; xyv[5 * n + 1] = y[n]; xyv[5 * n + 2] = vx[n]; xyv[5 * n + 3] = vy[n]; xyv[5 * n + 4] = vz[n];
    }
}

// ---------------------------------------------------------------------------------------------
// omega-k spectra of the field histories, kept on the device:
//   Electrostatic2D3V.jl:219,229   Z = log10.(sum(i -> abs.(fft(F[:, i, :])[2:end/2, 1:wind]), 1:size(F, 2)))'
//   PIC2D3V.jl:1483,1489           Z = log10.(abs.(fft(F)[2:kxind, 1, 1:wind]))'     (3D transform, k_perp = 0 slice)
// F is NA x NB x ND (column-major); the transform runs over (axis, time).  mode 0 ("sum of |.|"): for every line of
// the other axis, 2D FFT then abs, summed over the lines; mode 1 ("|.| of sum"): the k_perp = 0 slice of the 3D FFT,
// i.e. the 2D FFT of the sum over the other axis.  Both return |.| BEFORE the log10 and before the caller's slicing.
// One block transforms one (line, time) plane row by row in two passes through global scratch.
// ---------------------------------------------------------------------------------------------
struct ESSpecArgs {
    const double *F;      // history NA x NB x ND
    double2 *W;           // scratch [lines][n][ND]
    double *out;          // [n][ND] column-major (k fastest)
    const double2 *twn, *twt;
    int NA, NB, ND, axis, mode, n, lines, lgn, lgt;
};

// pass 1: forward along the chosen space axis for every (line, time); writes W[line][k][t] (k in natural order)
__global__ void __launch_bounds__(256) es_spec_space_kernel(ESSpecArgs a)
{
    extern __shared__ double smem[];
    const int n = a.n;
    double *re = smem, *im = smem + n;
    const int t = blockIdx.x, line = blockIdx.y;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double v;
        if (a.mode == 1) { // sum over the other axis first
            v = 0.0;
            const int other = a.axis == 0 ? a.NB : a.NA;
            for (int o = 0; o < other; ++o) {
                size_t idx = a.axis == 0 ? (size_t)k + (size_t)o * a.NA : (size_t)o + (size_t)k * a.NA;
                v += a.F[idx + (size_t)t * a.NA * a.NB];
            }
        } else {
            size_t idx = a.axis == 0 ? (size_t)k + (size_t)line * a.NA : (size_t)line + (size_t)k * a.NA;
            v = a.F[idx + (size_t)t * a.NA * a.NB];
        }
        re[k] = v; im[k] = 0.0;
    }
    __syncthreads();
    fft_smem4<false>(re, im, n, 1, 1, n, a.twn, n); // natural in, bit-reversed out
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        int kk = bitrev(k, a.lgn);
        a.W[((size_t)line * n + kk) * a.ND + t] = make_double2(re[k], im[k]);
    }
}

// pass 2: forward along time for every (line, k); out[k][w] += |.| (one atomic per element and line)
__global__ void __launch_bounds__(256) es_spec_time_kernel(ESSpecArgs a)
{
    extern __shared__ double smem[];
    const int ND = a.ND;
    double *re = smem, *im = smem + ND;
    const int k = blockIdx.x, line = blockIdx.y;
    const double2 *src = a.W + ((size_t)line * a.n + k) * ND;
    for (int t = threadIdx.x; t < ND; t += blockDim.x) { double2 z = src[t]; re[t] = z.x; im[t] = z.y; }
    __syncthreads();
    fft_smem4<false>(re, im, ND, 1, 1, ND, a.twt, ND);
    __syncthreads();
    for (int t = threadIdx.x; t < ND; t += blockDim.x) {
        int w = bitrev(t, a.lgt);
        atomicAdd(&a.out[(size_t)k + (size_t)w * a.n], sqrt(re[t] * re[t] + im[t] * im[t]));
    }
}

} // namespace pg
