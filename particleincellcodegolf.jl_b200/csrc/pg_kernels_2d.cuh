// pg_kernels_2d.cuh -- the fused particle loop of src/Electrostatic2D3V.jl:125-138:
//   gather E at the old position (eval :94-103) -> boris (:35-41) -> move + unimod (:131-132)
//   -> CIC deposit at the new position (depositcharge! :105-109).
// Five fp64 SoA streams in, five out (80 B per particle-step).  The field is kept as one
// (Ex,Ey) double2 per cell so a CIC corner is a single 16-byte load.  Charge accumulates in 64-bit fixed
// point (order-independent, bit-reproducible; see pg_kernels_1d.cuh).
#pragma once
#include "pg_common.cuh"
#include "pg_tma.cuh"

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
    fx_t *rho;           // global fixed-point deposit grid (sum of wx*wy; the factor w is applied by the solve)
    double *partials;    // [3*gridDim.x] per-block (sum vx^2+vy^2, sum vx, sum vy)
    long long P;
    double dt, t1, tscale;
    double fx_scale;     // 2^frac of the global grid
    int NX, NY;
    // tile-sorted mode
    const unsigned int *tile_start, *tile_end; // particle range of each tile in the sorted arrays
    const unsigned int *item_off;              // [ntiles+1] prefix of ceil(count/T2_CHUNK): work items per tile
    unsigned long long *slow_count;
    double fxw_scale;    // 2^frac of the shared-memory window (finer than the global grid)
    int fx_shift;        // window -> global: rounded right shift
    int ntx, ntiles;
};

// One particle: gather at the old position -> boris -> move -> CIC corners/weights at the new position.
struct Cic4 { int ix[2], iy[2]; double wx[2], wy[2]; };
__device__ __forceinline__ void cic4(double x, double y, int NX, int NY, Cic4 &c)
{
    cic_g(x, NX, c.ix[0], c.wx[0], c.ix[1], c.wx[1]);
    cic_g(y, NY, c.iy[0], c.wy[0], c.iy[1], c.wy[1]);
}

// Any particle order: E2 from global memory (L1/L2), deposit with 64-bit integer REDs to the global grid.
__global__ void __launch_bounds__(PG_THREADS) particles_2d3v_kernel(P2DArgs a)
{
    __shared__ double scratch[32];
    const int NX = a.NX, NY = a.NY;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.P; p += stride) {
        double x = ld_stream(a.x + p), y = ld_stream(a.y + p);
        double vx = ld_stream(a.vx + p), vy = ld_stream(a.vy + p), vz = ld_stream(a.vz + p);
        Cic4 c;
        cic4(x, y, NX, NY, c);
        double ex = 0.0, ey = 0.0;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
                double wxy = c.wx[ii] * c.wy[jj];
                double2 f = __ldg(&a.E2[(c.ix[ii] - 1) + (size_t)(c.iy[jj] - 1) * NX]);
                ex = fma(f.x, wxy, ex); // @muladd F1o += real(F1[i,j]) * wxy
                ey = fma(f.y, wxy, ey);
            }
        boris(vx, vy, vz, ex, ey, a.dt, a.t1, a.tscale);
        x = unimod(x + vx * a.dt, 1.0);
        y = unimod(y + vy * a.dt, 1.0);
        cic4(x, y, NX, NY, c);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) // F[i,j] += wx*wy*w
                atomicAdd(&a.rho[(c.ix[ii] - 1) + (size_t)(c.iy[jj] - 1) * NX], to_fx(c.wx[ii] * c.wy[jj], a.fx_scale));
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

// ------------------------------------------------------------------------------------------------
// Tile-sorted variant.  Particles are kept sorted by T2_TS x T2_TS-cell tile (pg_sort.cuh, mode 1).  A work
// item is up to T2_CHUNK consecutive particles of ONE tile; the block stages the (T2_TS+2*T2_R)^2 window of
// E2 around the tile in shared memory and accumulates the deposit into a fixed-point window of the same size
// (shared-memory 64-bit atomics), flushed once per item with integer REDs.  So the random-access traffic
// (4 field corners + 4 deposits per particle) never leaves the SM; HBM sees only the 80 B/particle streams.
// Particles that drifted more than T2_R cells out of their tile since the last sort use global memory.
// ------------------------------------------------------------------------------------------------
constexpr int T2_TS = 16, T2_SHIFT = 4, T2_R = 8, T2_WS = T2_TS + 2 * T2_R, T2_CHUNK = 8192;

// item_off[t] = sum_{u<t} ceil(count_u / T2_CHUNK); one block of 1024 threads, ntiles <= 4096.
__global__ void __launch_bounds__(1024) tile_worklist_kernel(const unsigned int *tile_start, const unsigned int *tile_end,
                                                            unsigned int *item_off, int ntiles)
{
    __shared__ unsigned int part[1024];
    const int per = (ntiles + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(lo + per, ntiles);
    unsigned int s = 0;
    for (int t = lo; t < hi; ++t) s += (tile_end[t] - tile_start[t] + T2_CHUNK - 1) / T2_CHUNK;
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        unsigned int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned int run = part[threadIdx.x] - s;
    for (int t = lo; t < hi; ++t) {
        item_off[t] = run;
        run += (tile_end[t] - tile_start[t] + T2_CHUNK - 1) / T2_CHUNK;
    }
    if (threadIdx.x == 1023) item_off[ntiles] = part[1023];
}

__global__ void __launch_bounds__(PG_THREADS) particles_2d3v_tiled(P2DArgs a)
{
    __shared__ double2 Ew[T2_WS * T2_WS];
    // fixed-point deposit window as two 32-bit limbs: 64-bit shared atomics are CAS loops on sm_100a
    // (ATOMS.CAST.SPIN.64), 32-bit adds are native; the carry out of the low limb is recovered from the
    // value the low-limb atomic returns.  hi cannot overflow: a cell receives at most T2_CHUNK units of weight.
    __shared__ unsigned int rlo[T2_WS * T2_WS], rhi[T2_WS * T2_WS];
    __shared__ double scratch[32];
    const int NX = a.NX, NY = a.NY, mx = NX - 1, my = NY - 1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
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
            Ew[c] = a.E2[gx + (size_t)gy * NX];
            rlo[c] = 0u; rhi[c] = 0u;
        }
        __syncthreads();
        for (long long p = start + threadIdx.x; p < end; p += blockDim.x) {
            double x = ld_stream(a.x + p), y = ld_stream(a.y + p);
            double vx = ld_stream(a.vx + p), vy = ld_stream(a.vy + p), vz = ld_stream(a.vz + p);
            Cic4 c;
            cic4(x, y, NX, NY, c);
            double ex = 0.0, ey = 0.0;
            {
                const int rx = (c.ix[0] - 1 - ox) & mx, ry = (c.iy[0] - 1 - oy) & my;
                if (rx <= T2_WS - 2 && ry <= T2_WS - 2) {
                    const double2 *e = Ew + rx + ry * T2_WS;
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                        for (int ii = 0; ii < 2; ++ii) {
                            double wxy = c.wx[ii] * c.wy[jj];
                            double2 f = e[ii + jj * T2_WS];
                            ex = fma(f.x, wxy, ex);
                            ey = fma(f.y, wxy, ey);
                        }
                } else {
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                        for (int ii = 0; ii < 2; ++ii) {
                            double wxy = c.wx[ii] * c.wy[jj];
                            double2 f = __ldg(&a.E2[(c.ix[ii] - 1) + (size_t)(c.iy[jj] - 1) * NX]);
                            ex = fma(f.x, wxy, ex);
                            ey = fma(f.y, wxy, ey);
                        }
                }
            }
            boris(vx, vy, vz, ex, ey, a.dt, a.t1, a.tscale);
            x = unimod(x + vx * a.dt, 1.0);
            y = unimod(y + vy * a.dt, 1.0);
            cic4(x, y, NX, NY, c);
            {
                const int rx = (c.ix[0] - 1 - ox) & mx, ry = (c.iy[0] - 1 - oy) & my;
                if (rx <= T2_WS - 2 && ry <= T2_WS - 2) {
                    const int r0 = rx + ry * T2_WS;
                    unsigned int lo[4], old[4];
                    fx_t v[4];
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                        for (int ii = 0; ii < 2; ++ii) {
                            const int k = ii + 2 * jj;
                            v[k] = to_fx(c.wx[ii] * c.wy[jj], a.fxw_scale);
                            lo[k] = (unsigned int)v[k];
                            old[k] = atomicAdd(&rlo[r0 + ii + jj * T2_WS], lo[k]);
                        }
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                        for (int ii = 0; ii < 2; ++ii) {
                            const int k = ii + 2 * jj;
                            const unsigned int carry = (old[k] + lo[k]) < old[k] ? 1u : 0u;
                            const unsigned int hi = (unsigned int)(v[k] >> 32) + carry;
                            if (hi) atomicAdd(&rhi[r0 + ii + jj * T2_WS], hi);
                        }
                } else {
                    ++nslow;
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                        for (int ii = 0; ii < 2; ++ii)
                            atomicAdd(&a.rho[(c.ix[ii] - 1) + (size_t)(c.iy[jj] - 1) * NX], to_fx(c.wx[ii] * c.wy[jj], a.fx_scale));
                }
            }
            st_stream(a.x + p, x); st_stream(a.y + p, y);
            st_stream(a.vx + p, vx); st_stream(a.vy + p, vy); st_stream(a.vz + p, vz);
            s0 += vx * vx + vy * vy; s1 += vx; s2 += vy;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < T2_WS * T2_WS; c += blockDim.x) {
            long long v = (long long)(((fx_t)rhi[c] << 32) | (fx_t)rlo[c]);
            if (v) {
                if (a.fx_shift > 0) v = (v + (1LL << (a.fx_shift - 1))) >> a.fx_shift; // rounded: weights are >= 0
                int gx = (ox + (c & (T2_WS - 1))) & mx, gy = (oy + (c >> 5)) & my;
                atomicAdd(&a.rho[gx + (size_t)gy * NX], (fx_t)v);
            }
        }
        __syncthreads();
    }
    s0 = block_sum(s0, scratch);
    s1 = block_sum(s1, scratch);
    s2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) {
        a.partials[3 * blockIdx.x] = s0; a.partials[3 * blockIdx.x + 1] = s1; a.partials[3 * blockIdx.x + 2] = s2;
    }
    if (nslow && a.slow_count) atomicAdd(a.slow_count, (unsigned long long)nslow);
}

// ------------------------------------------------------------------------------------------------
// Shared pieces of the streaming kernel below: the block's windows and one particle's work in two halves.
// ------------------------------------------------------------------------------------------------
struct P2DWindow {
    const double2 *Ew;
    unsigned int *rlo, *rhi;
    int ox, oy;
};

// One particle in two halves, so that a lane holding two particles can overlap their shared-memory latencies:
// p2d_push = gather (old position) -> boris -> move, returns the CIC corners/weights of the new position;
// p2d_deposit = the four window adds (or the global fallback).
// G, D: replicas per window cell of the field / of the deposit limbs (w.Ew, w.rlo, w.rhi already point at this lane's replica).
template <int G>
__device__ __forceinline__ void p2d_push(const P2DArgs &a, const P2DWindow &w, double &x, double &y, double &vx, double &vy,
                                         double &vz, double &s0, double &s1, double &s2, Cic4 &c)
{
    const int NX = a.NX, NY = a.NY, mx = NX - 1, my = NY - 1;
    cic4(x, y, NX, NY, c);
    double ex = 0.0, ey = 0.0;
    {
        const int rx = (c.ix[0] - 1 - w.ox) & mx, ry = (c.iy[0] - 1 - w.oy) & my;
        if (rx <= T2_WS - 2 && ry <= T2_WS - 2) {
            const double2 *e = w.Ew + (rx + ry * T2_WS) * G;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int ii = 0; ii < 2; ++ii) {
                    double wxy = c.wx[ii] * c.wy[jj];
                    double2 f = e[(ii + jj * T2_WS) * G];
                    ex = fma(f.x, wxy, ex);
                    ey = fma(f.y, wxy, ey);
                }
        } else {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int ii = 0; ii < 2; ++ii) {
                    double wxy = c.wx[ii] * c.wy[jj];
                    double2 f = __ldg(&a.E2[(c.ix[ii] - 1) + (size_t)(c.iy[jj] - 1) * NX]);
                    ex = fma(f.x, wxy, ex);
                    ey = fma(f.y, wxy, ey);
                }
        }
    }
    boris(vx, vy, vz, ex, ey, a.dt, a.t1, a.tscale);
    x = unimod(x + vx * a.dt, 1.0);
    y = unimod(y + vy * a.dt, 1.0);
    cic4(x, y, NX, NY, c);
    s0 += vx * vx + vy * vy; s1 += vx; s2 += vy;
}

template <int D>
__device__ __forceinline__ void p2d_deposit(const P2DArgs &a, const P2DWindow &w, const Cic4 &c, unsigned int &nslow)
{
    const int NX = a.NX, mx = NX - 1, my = a.NY - 1;
    const int rx = (c.ix[0] - 1 - w.ox) & mx, ry = (c.iy[0] - 1 - w.oy) & my;
    if (rx <= T2_WS - 2 && ry <= T2_WS - 2) {
        const int r0 = (rx + ry * T2_WS) * D;
        unsigned int lo[4], old[4];
        fx_t v[4];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
                const int k = ii + 2 * jj;
                v[k] = to_fx_small(c.wx[ii] * c.wy[jj], a.fxw_scale); // weights <= 1, fxw_scale <= 2^48
                lo[k] = (unsigned int)v[k];
                old[k] = atomicAdd(&w.rlo[r0 + (ii + jj * T2_WS) * D], lo[k]);
            }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
                const int k = ii + 2 * jj;
                const unsigned int carry = (old[k] + lo[k]) < old[k] ? 1u : 0u;
                const unsigned int hi = (unsigned int)(v[k] >> 32) + carry;
                if (hi) atomicAdd(&w.rhi[r0 + (ii + jj * T2_WS) * D], hi);
            }
    } else {
        ++nslow;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int ii = 0; ii < 2; ++ii)
                atomicAdd(&a.rho[(c.ix[ii] - 1) + (size_t)(c.iy[jj] - 1) * NX], to_fx(c.wx[ii] * c.wy[jj], a.fx_scale));
    }
}

template <int G, int D>
__device__ __forceinline__ void p2d_particle(const P2DArgs &a, const P2DWindow &w, double &x, double &y, double &vx, double &vy,
                                             double &vz, double &s0, double &s1, double &s2, unsigned int &nslow)
{
    Cic4 c;
    p2d_push<G>(a, w, x, y, vx, vy, vz, s0, s1, s2, c);
    p2d_deposit<D>(a, w, c, nslow);
}

// ------------------------------------------------------------------------------------------------
// Slice-streaming kernel (default of the tile-sorted path).  particles_2d3v_tiled above is bound by bytes in flight: its warps
// spend 10.8 of every 15.8 stall cycles per issue waiting for their five plain global loads (ncu, profiles/r1_d_*: DRAM 55 %,
// shared-memory pipe ~50 %).  A first rewrite that fed the same 8192-particle work items through per-warp cp.async rings
// (measured: 4.05 ms instead of 4.79, profiles/r2_e_*; since deleted) moved the limit to the shared-memory data pipe (80 %
// busy): per row of 32 particles 43 wavefronts for the four 16-byte field gathers (10.8 per LDS.128: eight lanes of a
// quarter-warp hit eight 16-byte bank groups at random), 41 for the eight limb atomics (5.1 per ATOMS) and 23 for the ring --
// 113 where 41 would do without bank conflicts.  This kernel streams the particles the same way and removes
// most of the conflicts by REPLICATING the windows: lane l reads the field from replica l % G (cell c, replica r at
// 16-byte slot c*G + r, so lanes of one quarter-warp collide only when they share r) and adds into deposit replica l % D
// (word c*D + r); the flush sums the D replicas of a cell.  The replicas fill the SM's shared memory (G = 4, D = 8: 64 + 64 KB
// for the 32 x 32-cell window), so ONE block of 512 threads runs per SM, and to keep it busy the work is no longer cut into
// 8192-particle items with a window rebuild each: a block owns a contiguous slice of the tile-sorted arrays (P / gridDim.x
// particles) and walks through it tile segment by tile segment -- at 2^28 particles 2-3 window builds per block and launch.
// The deposit window is flushed every S2_FLUSH particles (its high limb must not overflow) and at every tile change.
// Rows are 64 particles (two per lane, 16-byte cp.async.cg -> LDGSTS.128 that bypass L1, 128-bit stores), S2_STAGES per warp.
// ------------------------------------------------------------------------------------------------
constexpr int S2_FLUSH = 1 << 16, S2_STAGES = 2, S2_ROW = 5 * 32; // double2 slots per ring stage of one warp
__host__ __device__ constexpr size_t s2_smem_bytes(int G, int D, int threads)
{
    return (size_t)T2_WS * T2_WS * (16 * G + 8 * D) + 32 * 8 + (size_t)(threads / 32) * S2_STAGES * S2_ROW * 16;
}

template <int G, int D, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) particles_2d3v_stream(P2DArgs a)
{
    constexpr int NC = T2_WS * T2_WS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *Ew = reinterpret_cast<double2 *>(smem_raw);                  // [cell][G]
    unsigned int *rlo = reinterpret_cast<unsigned int *>(Ew + NC * G), *rhi = rlo + NC * D; // [cell][D]
    double *scratch = reinterpret_cast<double *>(rhi + NC * D);
    double2 *ring = reinterpret_cast<double2 *>(scratch + 32);            // [warp][stage][array][lane]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int nw = THREADS / 32;
    double2 *const mine = ring + (size_t)wid * S2_STAGES * S2_ROW + lane;
    const int NX = a.NX, NY = a.NY, mx = NX - 1, my = NY - 1;
    double *const gp[5] = {a.x, a.y, a.vx, a.vy, a.vz};
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    unsigned int nslow = 0;
    P2DWindow w;
    w.Ew = Ew + (lane & (G - 1)); w.rlo = rlo + (lane & (D - 1)); w.rhi = rhi + (lane & (D - 1)); w.ox = 0; w.oy = 0;
    for (int c = threadIdx.x; c < NC * D; c += THREADS) { rlo[c] = 0u; rhi[c] = 0u; }
    // this block's slice [s_lo, s_hi) of the sorted arrays (row aligned) and the tile its first particle lies in
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
        // rows of this segment: 64 particles from the even base; row k of this warp is row wid + k*nw
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
                Ew[c] = a.E2[gx + (size_t)gy * NX];
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
                if (v0 && v1) { // the two pushes are independent: their gathers overlap; then the eight + eight window adds
                    Cic4 c0, c1;
                    p2d_push<G>(a, w, X.x, Y.x, VX.x, VY.x, VZ.x, s0, s1, s2, c0);
                    p2d_push<G>(a, w, X.y, Y.y, VX.y, VY.y, VZ.y, s0, s1, s2, c1);
                    p2d_deposit<D>(a, w, c0, nslow);
                    p2d_deposit<D>(a, w, c1, nslow);
                } else if (v0) p2d_particle<G, D>(a, w, X.x, Y.x, VX.x, VY.x, VZ.x, s0, s1, s2, nslow);
                else p2d_particle<G, D>(a, w, X.y, Y.y, VX.y, VY.y, VZ.y, s0, s1, s2, nslow);
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
            if (v) {
                long long sv = (long long)v;
                if (a.fx_shift > 0) sv = (sv + (1LL << (a.fx_shift - 1))) >> a.fx_shift; // rounded: weights are >= 0
                int gx = (w.ox + (c & (T2_WS - 1))) & mx, gy = (w.oy + (c >> 5)) & my;
                atomicAdd(&a.rho[gx + (size_t)gy * NX], (fx_t)sv);
            }
        }
        pos = seg_hi; // the __syncthreads() at the top of the next segment orders the clears before its deposits
    }
    s0 = block_sum(s0, scratch);
    s1 = block_sum(s1, scratch);
    s2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) {
        a.partials[3 * blockIdx.x] = s0; a.partials[3 * blockIdx.x + 1] = s1; a.partials[3 * blockIdx.x + 2] = s2;
    }
    if (nslow && a.slow_count) atomicAdd(a.slow_count, (unsigned long long)nslow);
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
__global__ void stage_cic_deposit_kernel(const double *x, const double *y, long long count, int NX, int NY, double fx_scale,
                                         fx_t *rho)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    int ix[2], iy[2]; double wx[2], wy[2];
    cic_g(x[p], NX, ix[0], wx[0], ix[1], wx[1]);
    cic_g(y[p], NY, iy[0], wy[0], iy[1], wy[1]);
    for (int jj = 0; jj < 2; ++jj)
        for (int ii = 0; ii < 2; ++ii)
            atomicAdd(&rho[(ix[ii] - 1) + (size_t)(iy[jj] - 1) * NX], to_fx(wx[ii] * wy[jj], fx_scale));
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

// Exs[:,:,ti] .= real.(Ex); Eys[:,:,ti] .= real.(Ey); phis[:,:,ti] .= real.(pifft * phi)      src/Electrostatic2D3V.jl:171-173.
// phi holds the spectrum of the charge density with phi[1,1] = 0 (:143-144), so its inverse transform is rho - mean(rho);
// one block (runs on recorded steps only, and only when the handle keeps snapshots).
__global__ void __launch_bounds__(1024) snapshot2d_kernel(const double2 *E2, const double *rho, long long n, double *ex, double *ey, double *phi)
{
    __shared__ double scratch[32];
    __shared__ double mean_s;
    double s = 0.0;
    for (long long k = threadIdx.x; k < n; k += blockDim.x) s += rho[k];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) mean_s = s / (double)n;
    __syncthreads();
    const double mean = mean_s;
    for (long long k = threadIdx.x; k < n; k += blockDim.x) {
        const double2 e = E2[k];
        ex[k] = e.x; ey[k] = e.y; phi[k] = rho[k] - mean;
    }
}

} // namespace pg
