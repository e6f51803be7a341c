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
// TMA-staged variant of the tiled kernel (opt-in experiment, PICGOLF_2D_TMA=1: measured slower than
// particles_2d3v_tiled on B200 because the SM-side work, not the particle streams, limits it; see picgolf.cu).  Same tiles, same windows and
// the same per-particle arithmetic, but the five particle streams are moved by the copy engine: one persistent
// 512-thread block per SM owns a contiguous range of work items and streams them in sub-tiles of T2_SUB particles
// through a T2_STAGES-deep shared-memory ring (`cp.async.bulk` loads completing on mbarriers, in-place update,
// bulk stores).  Bulk copies need 16-byte alignment, so a sub-tile [a,b) is split into its even-aligned interior
// (staged) and at most two edge particles (plain global accesses).  The E / deposit window is kept across the
// consecutive items of a tile and flushed only when the tile changes.
// ------------------------------------------------------------------------------------------------
constexpr int T2_SUB = 1024, T2_STAGES = 3, T2_TMA_THREADS = 512;
constexpr size_t T2_TMA_SMEM = (size_t)T2_STAGES * 5 * T2_SUB * 8 + (size_t)T2_WS * T2_WS * (16 + 8) + 256 + 64;

struct P2DWindow {
    const double2 *Ew;
    unsigned int *rlo, *rhi;
    int ox, oy;
};

// gather (old position) -> boris -> move -> deposit (new position) for one particle; window or global fallback
__device__ __forceinline__ void p2d_particle(const P2DArgs &a, const P2DWindow &w, double &x, double &y, double &vx, double &vy,
                                             double &vz, double &s0, double &s1, double &s2, unsigned int &nslow)
{
    const int NX = a.NX, NY = a.NY, mx = NX - 1, my = NY - 1;
    Cic4 c;
    cic4(x, y, NX, NY, c);
    double ex = 0.0, ey = 0.0;
    {
        const int rx = (c.ix[0] - 1 - w.ox) & mx, ry = (c.iy[0] - 1 - w.oy) & my;
        const bool in = rx <= T2_WS - 2 && ry <= T2_WS - 2;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int ii = 0; ii < 2; ++ii) {
                double wxy = c.wx[ii] * c.wy[jj];
                double2 f = in ? w.Ew[rx + ii + (ry + jj) * T2_WS] : __ldg(&a.E2[(c.ix[ii] - 1) + (size_t)(c.iy[jj] - 1) * NX]);
                ex = fma(f.x, wxy, ex);
                ey = fma(f.y, wxy, ey);
            }
    }
    boris(vx, vy, vz, ex, ey, a.dt, a.t1, a.tscale);
    x = unimod(x + vx * a.dt, 1.0);
    y = unimod(y + vy * a.dt, 1.0);
    cic4(x, y, NX, NY, c);
    {
        const int rx = (c.ix[0] - 1 - w.ox) & mx, ry = (c.iy[0] - 1 - w.oy) & my;
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
                    old[k] = atomicAdd(&w.rlo[r0 + ii + jj * T2_WS], lo[k]);
                }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int ii = 0; ii < 2; ++ii) {
                    const int k = ii + 2 * jj;
                    const unsigned int carry = (old[k] + lo[k]) < old[k] ? 1u : 0u;
                    const unsigned int hi = (unsigned int)(v[k] >> 32) + carry;
                    if (hi) atomicAdd(&w.rhi[r0 + ii + jj * T2_WS], hi);
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
    s0 += vx * vx + vy * vy; s1 += vx; s2 += vy;
}

// Walks the sub-tiles of a contiguous range of work items (uniform across the block).
struct SubTileCursor {
    unsigned int item, item_end;
    int tile;
    long long pos, end; // next particle to hand out / end of the current item
    __device__ __forceinline__ void open_item(const P2DArgs &a)
    {
        int lo = 0, hi = a.ntiles;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (a.item_off[mid] <= item) lo = mid; else hi = mid;
        }
        tile = lo;
        pos = (long long)a.tile_start[tile] + (long long)(item - a.item_off[tile]) * T2_CHUNK;
        end = min(pos + (long long)T2_CHUNK, (long long)a.tile_end[tile]);
    }
    __device__ __forceinline__ void init(const P2DArgs &a, unsigned int i0, unsigned int i1)
    {
        item = i0; item_end = i1; tile = -1; pos = end = 0;
        if (item < item_end) open_item(a);
    }
    // next sub-tile [sa, sb) of tile t; false when the range is exhausted
    __device__ __forceinline__ bool next(const P2DArgs &a, long long &sa, long long &sb, int &t)
    {
        while (item < item_end && pos >= end) {
            ++item;
            if (item < item_end) open_item(a);
        }
        if (item >= item_end) return false;
        sa = pos; sb = min(pos + (long long)T2_SUB, end); t = tile;
        pos = sb;
        return true;
    }
};

__global__ void __launch_bounds__(T2_TMA_THREADS, 1) particles_2d3v_tma(P2DArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);                       // [STAGES][5][T2_SUB]
    double2 *Ew = reinterpret_cast<double2 *>(ring + (size_t)T2_STAGES * 5 * T2_SUB);
    unsigned int *rlo = reinterpret_cast<unsigned int *>(Ew + T2_WS * T2_WS), *rhi = rlo + T2_WS * T2_WS;
    double *scratch = reinterpret_cast<double *>(rhi + T2_WS * T2_WS);
    uint64_t *full = reinterpret_cast<uint64_t *>(scratch + 32);
    const int NX = a.NX, NY = a.NY, mx = NX - 1, my = NY - 1;
    double *const gp[5] = {a.x, a.y, a.vx, a.vy, a.vz};
    if (threadIdx.x == 0) {
        for (int s = 0; s < T2_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned int nitems = a.item_off[a.ntiles];
    const unsigned int i0 = (unsigned int)((unsigned long long)nitems * blockIdx.x / gridDim.x);
    const unsigned int i1 = (unsigned int)((unsigned long long)nitems * (blockIdx.x + 1) / gridDim.x);
    SubTileCursor prod, cons;
    prod.init(a, i0, i1);
    cons.init(a, i0, i1);
    long long issued = 0;
    auto issue_load = [&]() { // thread 0: next sub-tile of the producer cursor into stage issued % STAGES
        long long sa, sb; int t;
        if (!prod.next(a, sa, sb, t)) return;
        const int s = (int)(issued % T2_STAGES);
        const long long a2 = sa + (sa & 1), b2 = sb - (sb & 1);
        const uint32_t bytes = b2 > a2 ? (uint32_t)((b2 - a2) * 8) : 0u;
        mbar_arrive_expect_tx(&full[s], 5 * bytes);
        if (bytes)
            for (int q = 0; q < 5; ++q) bulk_load(ring + ((size_t)s * 5 + q) * T2_SUB, gp[q] + a2, bytes, &full[s]);
        ++issued;
    };
    if (threadIdx.x == 0)
        for (int k = 0; k < T2_STAGES - 1; ++k) issue_load();
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    unsigned int nslow = 0;
    P2DWindow w;
    w.Ew = Ew; w.rlo = rlo; w.rhi = rhi; w.ox = 0; w.oy = 0;
    int cur_tile = -1;
    auto flush_window = [&]() {
        for (int c = threadIdx.x; c < T2_WS * T2_WS; c += blockDim.x) {
            long long v = (long long)(((fx_t)rhi[c] << 32) | (fx_t)rlo[c]);
            if (v) {
                if (a.fx_shift > 0) v = (v + (1LL << (a.fx_shift - 1))) >> a.fx_shift;
                int gx = (w.ox + (c & (T2_WS - 1))) & mx, gy = (w.oy + (c >> 5)) & my;
                atomicAdd(&a.rho[gx + (size_t)gy * NX], (fx_t)v);
            }
        }
    };
    long long sa, sb, q = 0;
    int t;
    while (cons.next(a, sa, sb, t)) {
        if (t != cur_tile) { // (re)build the window: E2 slice in, deposit limbs flushed and cleared
            if (cur_tile >= 0) flush_window();
            __syncthreads();
            w.ox = (t % a.ntx) * T2_TS - T2_R; w.oy = (t / a.ntx) * T2_TS - T2_R;
            for (int c = threadIdx.x; c < T2_WS * T2_WS; c += blockDim.x) {
                int gx = (w.ox + (c & (T2_WS - 1))) & mx, gy = (w.oy + (c >> 5)) & my;
                Ew[c] = a.E2[gx + (size_t)gy * NX];
                rlo[c] = 0u; rhi[c] = 0u;
            }
            cur_tile = t;
            __syncthreads();
        }
        const int s = (int)(q % T2_STAGES);
        double *st = ring + (size_t)s * 5 * T2_SUB;
        const long long a2 = sa + (sa & 1), b2 = sb - (sb & 1);
        mbar_wait(&full[s], (uint32_t)((q / T2_STAGES) & 1));
        for (long long i = threadIdx.x; i < b2 - a2; i += blockDim.x) { // staged interior
            double x = st[i], y = st[T2_SUB + i], vx = st[2 * T2_SUB + i], vy = st[3 * T2_SUB + i], vz = st[4 * T2_SUB + i];
            p2d_particle(a, w, x, y, vx, vy, vz, s0, s1, s2, nslow);
            st[i] = x; st[T2_SUB + i] = y; st[2 * T2_SUB + i] = vx; st[3 * T2_SUB + i] = vy; st[4 * T2_SUB + i] = vz;
        }
        // unaligned edge particles straight from/to global memory (threads of two different warps)
        long long edge = -1; // an odd start and an odd end can never be the same particle
        if (threadIdx.x == 0 && (sa & 1)) edge = sa;
        if (threadIdx.x == 32 && (sb & 1)) edge = sb - 1;
        if (edge >= 0) {
            double x = a.x[edge], y = a.y[edge], vx = a.vx[edge], vy = a.vy[edge], vz = a.vz[edge];
            p2d_particle(a, w, x, y, vx, vy, vz, s0, s1, s2, nslow);
            a.x[edge] = x; a.y[edge] = y; a.vx[edge] = vx; a.vy[edge] = vy; a.vz[edge] = vz;
        }
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            if (b2 > a2) {
                const uint32_t bytes = (uint32_t)((b2 - a2) * 8);
                for (int k = 0; k < 5; ++k) bulk_store(gp[k] + a2, st + (size_t)k * T2_SUB, bytes);
            }
            bulk_commit();
            bulk_wait_read<1>(); // the stage stored one iteration ago is free again
            issue_load();
        }
        ++q;
    }
    if (cur_tile >= 0) { __syncthreads(); flush_window(); }
    if (threadIdx.x == 0) bulk_wait_all<0>();
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
