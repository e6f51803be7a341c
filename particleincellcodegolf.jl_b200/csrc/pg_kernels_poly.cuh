// pg_kernels_poly.cuh -- cell-polynomial form of the Gaussian fixed-point pass (src/GaussianFixedPoint.jl:6-9,
// src/GaussianFixedPointQuiet.jl:7-10) for shards with many particles per cell.
//
// For a power-of-two grid every weight of the stencil d(c) is a polynomial in t = 2*delta, delta = c*N - round(c*N)
// (gauss_cellpoly.inc: W_j = sum_n PG_CW[j+6][n] t^n, degree 16, |error| <= 7e-17).  Both particle<->grid maps of a
// sweep are linear in the weights, so they factor through the powers of t:
//     gather   sum(k->E[k[1]]*k[2], d(c))  = sum_n t^n G[cell][n],   G[cell][n] = sum_j CW[j][n] E[cell+j]   gpoly_kernel
//     deposit  r[k[1]] += k[2]             : M[cell][n] += t^n,       rho[i] = sum_j sum_n CW[j][n] M[i-j][n]  mom2rho_kernel
// A particle-sweep then costs one Horner evaluation (16 DFMA) and 16 power accumulations instead of two 13-weight
// stencil evaluations (~255 FP64 instructions in fp_pass_sorted), which moves the pass from the FP64 pipe to HBM.
//
// fp_pass_poly: every warp streams ONE contiguous range of the (cell, sign v)-sorted particle arrays with 128-bit
// loads (2 particles per lane and row).  Deposit: each lane keeps the moment set of the cell it is currently in in
// registers and a spare set -- the other cell of the drifting bin, which straddles two neighbouring cells -- in a
// private shared-memory column; a set is flushed with 17 integer REDs into the fixed-point moment grid Mg only when
// the lane meets a third cell -- a few times per pass in sorted order, so there are no shared-memory or per-particle
// atomics at all.
// Gather: the G rows of the CP_WG cells around the warp's position are staged in shared memory; a particle reads the
// 17 coefficients of its own cell (lanes in the same cell broadcast).  Any particle order is handled correctly
// (window reloads, global-memory gather, early flushes); order only decides the speed, and the mid-stream flushes
// are counted so that the host can re-sort sooner when lanes start to alternate between cells (picgolf_sort_stats).
#pragma once
#include "pg_kernels_1d.cuh"
#include "gauss_cellpoly.inc"

namespace pg {

constexpr int CP_NC = PG_CW_NC; // coefficients / moments per cell (degree 16)
constexpr int CP_NM = CP_NC - 1; // moments kept as doubles (n = 1..16); n = 0 is an integer count
constexpr int CP_WG = 8;         // cells in a warp's gather window
constexpr int CP_GS = CP_NC + 1; // row stride of the gather tables: coefficient pairs (c_2m, c_2m+1) are 16-byte aligned,
                                 // and two neighbouring rows (144 B apart) never share a bank within one 128-bit access

// 3 blocks of 128 threads per SM (160 registers per thread): 4 x 128 at 128 registers measured the same, 5 x 128 at 96
// registers 19 % slower -- the compiler needs the registers to overlap the coefficient loads of a lane's two particles.
#ifndef PG_CP_THREADS
#define PG_CP_THREADS 128
#endif
#ifndef PG_CP_MINBLOCKS
#define PG_CP_MINBLOCKS 3
#endif
constexpr int CP_THREADS = PG_CP_THREADS; // threads per block of fp_pass_poly
constexpr int CP_STAGES = 4;     // rows of particle data in flight per warp (cp.async ring, 1.5 KB per stage)
constexpr int CP_STAGE_D2 = 96;  // double2 slots per stage: X, V, v pairs of the 32 lanes

__host__ __device__ inline size_t cp_smem_bytes(int threads)
{
    return (size_t)(threads / 32) * (CP_GS * CP_WG * sizeof(double) + (size_t)CP_STAGES * CP_STAGE_D2 * sizeof(double2)) +
           (size_t)threads * CP_NM * sizeof(double); // + the spare moment set of every lane
}

// 16-byte asynchronous global -> shared copy (L2 only); each lane later reads back exactly the bytes it copied itself,
// so cp.async.wait_group alone orders the accesses.
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// G[z*CP_GS + n] = sum_j CW[j][n] * E[(z+j) mod N]  for the 0-based cell z = mod1(round(c*N),N)-1; also clears the moment
// grid the previous mom2rho_kernel consumed.  Skipped (like the solve) once the step has converged.
struct GPolyArgs {
    const double *E;
    double *G;   // [N][CP_GS]
    fx_t *Mg;    // [N][CP_NC]
    const Ctrl *ctrl;
    int N, k;
};

__global__ void __launch_bounds__(128) gpoly_kernel(GPolyArgs a)
{
    const int fk = a.ctrl->final_k;
    if (fk >= 0 && a.k > fk) return;
    const int N = a.N, Nmask = N - 1;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    if (z < N) {
        double e[GAUSS_NW];
#pragma unroll
        for (int q = 0; q < GAUSS_NW; ++q) e[q] = a.E[(z + q - 6) & Nmask];
#pragma unroll
        for (int n = 0; n < CP_NC; ++n) {
            double g = 0.0;
#pragma unroll
            for (int q = 0; q < GAUSS_NW; ++q) g = fma(PG_CW[q][n], e[q], g);
            a.G[(size_t)z * CP_GS + n] = g;
        }
        a.G[(size_t)z * CP_GS + CP_NC] = 0.0;
    }
    const int lo = blockIdx.x * blockDim.x * CP_NC, hi = min(N * CP_NC, lo + (int)blockDim.x * CP_NC);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) a.Mg[i] = 0ULL;
}

// rho_fx[i] += fixed( sum_j sum_n CW[j][n] * M[(i-j) mod N][n] ).  One block = CPM_CELLS output cells; the moments of
// the CPM_CELLS+12 source cells are converted to fp64 once in shared memory; 4 threads share the 13 offsets of a cell
// and are summed in a fixed order.  Mg is cleared later by gpoly_kernel (other blocks read the halo cells).
struct Mom2RhoArgs {
    const fx_t *Mg;
    fx_t *rho;
    const Ctrl *ctrl;
    double fx_scale, fx_inv;
    int N;
    const unsigned long long *flush_src; // NCCL path on several GPUs: this rank's flush counter goes to rho[N] and is summed with the grid
};
constexpr int CPM_CELLS = 64;

__global__ void __launch_bounds__(4 * CPM_CELLS) mom2rho_kernel(Mom2RhoArgs a)
{
    __shared__ double Ms[(CPM_CELLS + 12) * CP_NC];
    __shared__ double part[4][CPM_CELLS];
    if (a.ctrl->final_k >= 0) return; // converged: the moments belong to the next step's first solve
    const int N = a.N, Nmask = N - 1;
    const int i0 = blockIdx.x * CPM_CELLS;
    if (a.flush_src && blockIdx.x == 0 && threadIdx.x == 0) a.rho[N] = *a.flush_src;
    for (int t = threadIdx.x; t < (CPM_CELLS + 12) * CP_NC; t += blockDim.x) {
        const int c = t / CP_NC, n = t - c * CP_NC;
        Ms[t] = (double)(long long)a.Mg[(size_t)((i0 + c - 6) & Nmask) * CP_NC + n] * a.fx_inv;
    }
    __syncthreads();
    const int i = threadIdx.x & (CPM_CELLS - 1), p = threadIdx.x / CPM_CELLS; // offsets q = p, p+4, p+8, (p+12)
    double s = 0.0;
    for (int q = p; q < GAUSS_NW; q += 4) {
        // cell i receives W_j from source cell i - j, j = q - 6; source slot = (i - j) + 6 = i + 12 - q
        const double *m = Ms + (i + 12 - q) * CP_NC;
#pragma unroll
        for (int n = 0; n < CP_NC; ++n) s = fma(PG_CW[q][n], m[n], s);
    }
    part[p][i] = s;
    __syncthreads();
    if (threadIdx.x < CPM_CELLS && i0 + i < N) {
        const double r = (part[0][i] + part[1][i]) + (part[2][i] + part[3][i]);
        a.rho[i0 + i] += to_fx(r, a.fx_scale);
    }
}

// Centre cell Int(round(c*N)) and t = 2*(c*N - round(c*N)) of a stencil at the midpoint c = (x+X)/2, from the sum
// s = x+X.  N is a power of two, so s*N = 2*c*N exactly; adding 1.5*2^53 (ulp 2) rounds it to the nearest EVEN integer,
// ties to the even multiple -- exactly 2*rint(c*N) for |c*N| < 2^51 -- and leaves rint(c*N) in the low mantissa word.
// Same bits as the literal  cn = ((x+X)/2)*N; r = rint(cn); t = 2*(cn-r), without the conversion unit and 3 FP64
// instructions shorter.
__device__ __forceinline__ void cp_centre(double s, double dN, int &cell, double &t)
{
    const double cn2 = s * dN;
    const double big = cn2 + 13510798882111488.0;
    const double r2 = big - 13510798882111488.0;
    cell = __double2loint(big);
    t = cn2 - r2;
}

// One lane's moment set for the cell it is currently in (n = 0 is an integer count).
struct CPSet {
    double m[CP_NM];
    int cnt, cell;
};

// Add the set to the fixed-point moment grid (17 integer REDs) and clear it.
__device__ __forceinline__ void cp_flush(CPSet &s, fx_t *Mg, double fx_scale, int Nmask)
{
    if (s.cnt) {
        fx_t *p = Mg + (size_t)((s.cell - 1) & Nmask) * CP_NC; // Julia index -> 0-based cell
        atomicAdd(p, to_fx((double)s.cnt, fx_scale));
#pragma unroll
        for (int n = 0; n < CP_NM; ++n) { atomicAdd(p + 1 + n, to_fx(s.m[n], fx_scale)); s.m[n] = 0.0; }
        s.cnt = 0;
    }
}

// The moment set of the lane's CURRENT cell lives in registers (CPSet); the set of the other cell of the drifting bin
// (the "spare") in a lane-private shared-memory column col[n*blockDim.x] (conflict-free), exchanged with the register
// set when the lane changes over.  (Both sets in registers, selected by cell parity: 4 % slower, 32 registers more;
// a branch-free sum / odd-sum form: another 4 % slower.  Both measured on B200.)
struct CPSpare { int cnt, cell; };

__device__ __forceinline__ void cp_flush_spare(CPSpare &sp, double *col, int stride, fx_t *Mg, double fx_scale, int Nmask)
{
    if (sp.cnt) {
        fx_t *p = Mg + (size_t)((sp.cell - 1) & Nmask) * CP_NC;
        atomicAdd(p, to_fx((double)sp.cnt, fx_scale));
#pragma unroll
        for (int n = 0; n < CP_NM; ++n) { atomicAdd(p + 1 + n, to_fx(col[n * stride], fx_scale)); col[n * stride] = 0.0; }
        sp.cnt = 0;
    }
}

__device__ __forceinline__ void cp_deposit1(int cell, double t, CPSet &P, CPSpare &sp, double *col, int stride, fx_t *Mg,
                                            double fx_scale, int Nmask, unsigned int &nflush)
{
    if (cell != P.cell) { // rare in sorted order
        if (cell == sp.cell) { // change over to the other cell of the bin: exchange the two sets
#pragma unroll
            for (int n = 0; n < CP_NM; ++n) { const double m = col[n * stride]; col[n * stride] = P.m[n]; P.m[n] = m; }
            const int c = sp.cnt; sp.cnt = P.cnt; P.cnt = c;
            sp.cell = P.cell; P.cell = cell;
            ++nflush; // exchanges count as well: lanes alternating between the two cells pay 32 shared accesses each time
        } else {               // a third cell: retire the spare, park the current set
            nflush += sp.cnt > 0;
            cp_flush_spare(sp, col, stride, Mg, fx_scale, Nmask);
#pragma unroll
            for (int n = 0; n < CP_NM; ++n) { col[n * stride] = P.m[n]; P.m[n] = 0.0; }
            sp.cnt = P.cnt; sp.cell = P.cell;
            P.cnt = 0; P.cell = cell;
        }
    }
    P.cnt++;
    const double t2 = t * t;
    double pa = t, pb = t2;
#pragma unroll
    for (int n = 0; n < CP_NM; n += 2) {
        P.m[n] += pa; P.m[n + 1] += pb;
        if (n + 2 < CP_NM) { pa *= t2; pb *= t2; }
    }
}

// Rare path: the particle's cell is outside the warp's gather window.
__device__ __noinline__ double cp_slow_gather(const double *G, int cell, double t, int N)
{
    const double *g = G + (size_t)((cell - 1) & (N - 1)) * CP_GS;
    const double t2 = t * t; // same operation order as cp_horner
    double ge = g[16], go = g[15];
    for (int n = 14; n >= 0; n -= 2) ge = fma(ge, t2, g[n]);
    for (int n = 13; n >= 1; n -= 2) go = fma(go, t2, g[n]);
    return fma(go, t, ge);
}

// sum_n t^n g[n] from one staged row (128-bit shared loads of the pairs (c_2m, c_2m+1)): even and odd halves as two
// independent Horner chains in t^2.
__device__ __forceinline__ double cp_horner(const double *g, double t)
{
    const double2 *g2 = reinterpret_cast<const double2 *>(g);
    const double t2 = t * t;
    double2 c = g2[8];
    double ge = c.x; // c_16
    c = g2[7];
    ge = fma(ge, t2, c.x);
    double go = c.y; // c_15
#pragma unroll
    for (int m = 6; m >= 0; --m) {
        c = g2[m];
        ge = fma(ge, t2, c.x);
        go = fma(go, t2, c.y);
    }
    return fma(go, t, ge);
}

// The last P % 64 particles of a shard (no full row): one thread each, global-memory gather polynomial and direct
// moment REDs.  Same arithmetic as the streaming loop below.
template <bool FIRST>
__device__ __forceinline__ void cp_tail_particle(const FPArgs &a, long long j, bool final, bool v0_is_V, double &sv2, double &sv)
{
    const int N = a.N, Nmask = N - 1;
    const double dN = (double)N, dt = a.dt;
    double Xj = a.X[j], Vj = a.V[j], vj = v0_is_V ? Vj : a.v[j];
    double xj = Xj + (vj + Vj) / 2 * dt;
    if (!FIRST) {
        const double cn = ((xj + Xj) / 2) * dN, rr = rint(cn), d = cn - rr;
        vj = Vj + cp_slow_gather(a.G, (int)rr, d + d, N) * dt;
        a.v[j] = vj;
        if (final) {
            const double xw = jl_mod1(xj);
            a.xout[j] = xw;
            sv2 = fma(vj, vj, sv2); sv += vj;
            Xj = xw; Vj = vj;
        }
        xj = Xj + (vj + Vj) / 2 * dt;
    }
    const double cn = ((xj + Xj) / 2) * dN, rr = rint(cn), d = cn - rr, t = d + d;
    fx_t *p = a.Mg + (size_t)(((int)rr - 1) & Nmask) * CP_NC;
    double pw = 1.0;
    for (int n = 0; n < CP_NC; ++n) { atomicAdd(p + n, to_fx(pw, a.fx_scale)); pw *= t; }
}

// Pass k of a step (same contract as fp_pass_sorted / fp_pass_atomic; FPArgs.G / FPArgs.Mg carry the polynomial
// tables).  Row = 64 consecutive particles; lane l owns particles 2l and 2l+1 of the row.
template <bool FIRST>
__global__ void __launch_bounds__(CP_THREADS, PG_CP_MINBLOCKS) fp_pass_poly(FPArgs a)
{
    extern __shared__ double smem[];
    __shared__ double scratch[32];
    const int fk = a.ctrl->final_k;
    if (!FIRST && fk >= 0 && a.k > fk) return;
    const bool final = !FIRST && fk == a.k;
    const bool v0_is_V = FIRST || a.k == 1; // sweep 1 starts from v = V: the work buffer is stale until pass 1 writes it
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double *Gw = smem + warp * (CP_GS * CP_WG); // [CP_WG][CP_GS]
    const int N = a.N, Nmask = N - 1;
    const double dN = a.dN, hdt = a.dt / 2; // ((v+V)/2)*dt == (v+V)*(dt/2) bit for bit (exact power-of-two scalings)
    // full rows only; every warp streams one contiguous range [r0, r1)
    const int rows = (int)(a.P >> 6);
    const int nw = gridDim.x * wpb, gw = blockIdx.x * wpb + warp;
    const int rpw = (rows + nw - 1) / nw;
    const int r0 = min(rows, gw * rpw), r1 = min(rows, r0 + rpw);
    const double2 *X2 = reinterpret_cast<const double2 *>(a.X), *V2 = reinterpret_cast<const double2 *>(a.V);
    double2 *v2 = reinterpret_cast<double2 *>(a.v), *xo2 = reinterpret_cast<double2 *>(a.xout);
    CPSet A;
    CPSpare spare;
    double *col = smem + wpb * (CP_GS * CP_WG + 2 * CP_STAGES * CP_STAGE_D2) + threadIdx.x;
    const int cstride = blockDim.x;
#pragma unroll
    for (int n = 0; n < CP_NM; ++n) { A.m[n] = 0.0; col[n * cstride] = 0.0; }
    A.cnt = 0; A.cell = 0x40000000; spare.cnt = 0; spare.cell = 0x40000001;
    int gb = 0x40000000; // Julia index of window slot 0; the first row always restages (see `staged`)
    bool staged = false;
    double sv2 = 0.0, sv = 0.0;
    unsigned int nflush = 0;
    // particle rows arrive through a per-warp cp.async ring: no registers are held while a row is in flight
    double2 *ring = reinterpret_cast<double2 *>(smem + wpb * (CP_GS * CP_WG)) + warp * (CP_STAGES * CP_STAGE_D2) + lane;
    long long j2 = ((long long)r0 << 5) + lane; // double2 index of this lane's pair
    auto issue = [&](int r, long long jj) {     // row r -> stage r % CP_STAGES (always commits: uniform group count)
        if (r < r1) {
            double2 *st = ring + ((r - r0) % CP_STAGES) * CP_STAGE_D2;
            cp_async16(st, X2 + jj);
            cp_async16(st + 32, V2 + jj);
            if (!v0_is_V) cp_async16(st + 64, v2 + jj);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < CP_STAGES - 1; ++s) issue(r0 + s, j2 + 32 * s);
    int stage = 0;
    for (int r = r0; r < r1; ++r, j2 += 32) {
        issue(r + CP_STAGES - 1, j2 + 32 * (CP_STAGES - 1));
        cp_async_wait<CP_STAGES - 1>(); // row r has landed
        const double2 *st = ring + stage * CP_STAGE_D2;
        stage = stage + 1 == CP_STAGES ? 0 : stage + 1;
        const double2 Xc = st[0], Vc = st[32];
        const double2 vc = v0_is_V ? Vc : st[64];
        double Xj[2] = {Xc.x, Xc.y}, Vj[2] = {Vc.x, Vc.y}, vj[2] = {vc.x, vc.y}, xj[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) xj[q] = Xj[q] + (vj[q] + Vj[q]) * hdt; // x.=X.+(v.+V)/2*dt
        if (!FIRST) {
            int cell[2];
            double t[2];
            unsigned int slot[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                cp_centre(xj[q] + Xj[q], dN, cell[q], t[q]);
                slot[q] = (unsigned int)(cell[q] - gb) & (unsigned int)Nmask;
            }
            if (!__all_sync(0xffffffffu, staged && slot[0] < CP_WG && slot[1] < CP_WG)) {
                // recentre the window one cell below the smallest centre of this row and restage it
                int cm = min(cell[0], cell[1]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cm = min(cm, __shfl_xor_sync(0xffffffffu, cm, o));
                gb = cm - 1;
                staged = true;
                __syncwarp();
                for (int i = lane; i < CP_GS * CP_WG; i += 32) {
                    const int s = i / CP_GS, n = i - s * CP_GS;
                    Gw[i] = a.G[(size_t)((gb + s - 1) & Nmask) * CP_GS + n];
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 2; ++q) slot[q] = (unsigned int)(cell[q] - gb) & (unsigned int)Nmask;
            }
            double g[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) g[q] = cp_horner(Gw + (slot[q] < CP_WG ? slot[q] : 0u) * CP_GS, t[q]);
            if (slot[0] >= CP_WG || slot[1] >= CP_WG) { // rare: a centre more than CP_WG cells above the row's smallest
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (slot[q] >= CP_WG) g[q] = cp_slow_gather(a.G, cell[q], t[q], N);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) vj[q] = Vj[q] + g[q] * a.dt; // v[j]=V[j]+sum(...)*dt
            __stcs(v2 + j2, make_double2(vj[0], vj[1]));
            if (final) {
                // end of step: x.=mod.(x,1), diagnostics sums -- and the first pass of the NEXT step fused in
                // (X.=x; V.=v; x = X + (V+V)/2*dt; deposit at (x+X)/2)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    Xj[q] = jl_mod1(xj[q]);
                    sv2 = fma(vj[q], vj[q], sv2); sv += vj[q];
                    Vj[q] = vj[q];
                }
                __stcs(xo2 + j2, make_double2(Xj[0], Xj[1]));
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) xj[q] = Xj[q] + (vj[q] + Vj[q]) * hdt;
        }
        {
            int cell[2];
            double t[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) cp_centre(xj[q] + Xj[q], dN, cell[q], t[q]);
            cp_deposit1(cell[0], t[0], A, spare, col, cstride, a.Mg, a.fx_scale, Nmask, nflush);
            cp_deposit1(cell[1], t[1], A, spare, col, cstride, a.Mg, a.fx_scale, Nmask, nflush);
        }
    }
    cp_async_wait<0>();
    cp_flush(A, a.Mg, a.fx_scale, Nmask);
    cp_flush_spare(spare, col, cstride, a.Mg, a.fx_scale, Nmask);
    // ragged tail of the shard: fewer than 64 particles, first warp of the last block
    if (blockIdx.x == gridDim.x - 1 && warp == 0) {
        const long long j = ((long long)rows << 6) + lane;
        if (j < a.P) cp_tail_particle<FIRST>(a, j, final, v0_is_V, sv2, sv);
        if (j + 32 < a.P) cp_tail_particle<FIRST>(a, j + 32, final, v0_is_V, sv2, sv);
    }
    if (final) {
        sv2 = block_sum(sv2, scratch);
        sv = block_sum(sv, scratch);
        if (threadIdx.x == 0) { a.partials[2 * blockIdx.x] = sv2; a.partials[2 * blockIdx.x + 1] = sv; }
    }
    if (nflush && a.slow_count) atomicAdd(a.slow_count, (unsigned long long)nflush);
}

} // namespace pg
