// pg_kernels_poly.cuh -- sub-cell polynomial form of the Gaussian fixed-point pass (src/GaussianFixedPoint.jl:6-9,
// src/GaussianFixedPointQuiet.jl:7-10) for shards with many particles per cell.
//
// For a power-of-two grid every weight of the stencil d(c) is a function of delta = c*N - round(c*N) only.  A cell is cut
// into CP_NSUB = 8 intervals: y = c*N*8 (exact: a power-of-two scaling of x+X), interval m = round(y), u = y - m.  On the
// sub-interval s = (m+4)&7 of stencil centre k = (m+4)>>3 each weight is a degree-10 polynomial in u
// (gauss_cellpoly.inc: W_j = sum_n PG_CWS[s][j+6][n] u^n, |error| <= 9e-17 for |u| <= 1, i.e. twice the interval).
// Both particle<->grid maps of a sweep are linear in the weights, so they factor through the powers of u:
//     gather   sum(k->E[k[1]]*k[2], d(c))  = sum_n u^n G[m][n],    G[m][n] = sum_j CWS[s][j][n] E[k+j]              gpoly_kernel
//     deposit  r[k[1]] += k[2]             : M[m][n] += u^n,        rho[i] = sum_j sum_s sum_n CWS[s][j][n] M[m(i-j,s)][n]   mom2rho_kernel
// A particle-sweep then costs one Horner evaluation (10 DFMA) and 10 power accumulations -- 47 FP64 instructions instead
// of the ~255 of the two 13-weight stencil evaluations in fp_pass_sorted, and of the 66 of the one-polynomial-per-cell
// form (degree 16) that round 1 shipped; that moves the pass from the FP64 pipe to HBM.  (16 intervals of degree 8 fitted on
// |u| <= 7/8 -- 41 instructions, -DPG_CWS_ALT -- are 2.7 % faster per step on cold beams and 10 - 28 % slower on warm ones,
// where the narrower intervals turn more deposits into outliers; the constants below follow the table.)
//
// fp_pass_poly: every warp streams ONE contiguous range of the (cell, sign v, sub-cell position)-sorted particle arrays
// with 128-bit loads (2 particles per lane and row).  Deposit: the warp shares ONE current interval; each lane sums u^n of its
// particles in registers, the 32 lane sets are added up and flushed with 11 integer REDs into the fixed-point moment grid Mg
// when the stream has moved on to the next interval (cp_deposit_row) -- no shared-memory atomics and no per-particle atomics
// at all.  (Deterministic mode: one set per lane with a spare, see below.)
// Gather: the G rows of the CP_WG intervals around the warp's position are staged in shared memory; a particle reads the
// 11 coefficients of its own interval (lanes in the same interval broadcast).  Any particle order is handled correctly
// (window reloads, global-memory gather, early flushes); order only decides the speed, and the flushes are counted so
// that the host can re-sort sooner when the bins shear apart (picgolf_sort_stats).
//
// DET = true (picgolf_config.deterministic = 1): the same pass, bit-reproducible from run to run.  The counting sort ranks
// the particles of a bin with atomics, so WHICH lane sums which particles changes between runs.  Two things follow.  The
// interval a particle deposits into must not depend on the lane's history: no hysteresis, always m = round(y); a lane
// that sits on an interval edge then alternates between two sets, so the second one (the "spare") lives in a lane-private
// shared-memory column and is exchanged with the register set on a change-over (a third interval retires the spare).
// And the lane sums integers: u^n + 1.5*2^2 has the fixed-point value of u^n (50 fractional bits) in its low mantissa bits, and the raw 64-bit
// patterns are added with integer adds (the offset, count * bits(6.0), is taken off at the flush) -- exact, associative,
// order-free, and more accurate than a running fp64 sum.  A set is flushed as two 64-bit words per moment (low 32 bits,
// high part) so the grid-wide integer sums cannot overflow; mom2rho_kernel puts them together.  The diagnostics sums
// (sum v^2, sum v) use the same trick with 40 fractional bits.
//
// Re-sort fused into the passes (FPArgs.fs_hist != NULL, one step at a time, chosen by the host): the particles are kept in the
// order of the NEXT step's first mid-point x + v*dt/2, binned by (cell centre, sign v, sub-cell position) -- the position every
// sweep of that step deposits and gathers at, up to E*dt^2/4, whatever the velocity spread.  A final pass k writes
// x = X + (v_{k-1}+V)/2*dt and v_k, and x depends on the OUTPUT of pass k-1 only, so the bin of (x, v_{k-1}) is known exactly one
// pass early: every pass k >= 1 that is not final counts these bins (runs of equal keys in a warp row share one RED);
// cp_fs_scan_kernel turns the counts into slot cursors if the solve declared sweep k+1 final (and clears them otherwise); the
// final pass recomputes the same key from the same operands (bit-identical: -fmad=false, the key's own fma is explicit),
// reserves its slots from the cursors (one atomic per run) and writes x, v and the particle's original index there -- into
// the other halves of the ping-pong buffers (v: a third buffer, the work buffer is read in place), so the sort adds no pass over the particles:
// 8 B of ids per particle and scattered instead of streaming stores.  A step that ends at sweep 1 has no counting pass before it
// and writes to the same buffers unpermuted, so the host's buffer rotation never depends on the sweep count.
#pragma once
#include "pg_kernels_1d.cuh"
#ifdef PG_CWS_ALT // A/B builds: 16 intervals per cell, degree 8, fitted on |u| <= 7/8 (tools/gen_gauss_cellpoly.py --nsub 16 --deg 8 --range 0.875)
#include "gauss_cellpoly_16_8.inc"
#else
#include "gauss_cellpoly.inc"
#endif
#include <type_traits>

namespace pg {

constexpr int CP_NSUB = PG_CWS_NSUB; // polynomial intervals per cell
constexpr int CP_SUBLG = CP_NSUB == 16 ? 4 : 3;
static_assert((1 << CP_SUBLG) == CP_NSUB, "CP_NSUB must be 2^CP_SUBLG");
constexpr int CP_NC = PG_CWS_NC; // coefficients / moments per interval (degree 10)
constexpr int CP_NM = CP_NC - 1; // moments kept as doubles (n = 1..10); n = 0 is an integer count
constexpr int CP_WG = 2 * CP_NSUB; // intervals in a warp's gather window (two cells)
constexpr int CP_GS = CP_NC + 1; // row stride of the gather tables: coefficient pairs (c_2m, c_2m+1) are 16-byte aligned
constexpr double CP_UMAX = PG_CWS_UMAX; // the polynomials are fitted on |u| <= CP_UMAX: a particle stays with the warp's interval that far from its centre
constexpr double CP_UMOVE = CP_UMAX - 0.25; // default mode: the warp moves on when most of a row has |u| > CP_UMOVE (after the move those particles sit
                                            // within 1/4 of the new centre and the rest of the row within CP_UMOVE: the warp never moves straight back)
constexpr double CP_MAGIC = 6755399441055744.0; // 1.5 * 2^52: y + CP_MAGIC rounds y to the nearest integer (ties to even), |y| < 2^51

#ifndef PG_CP_THREADS
#define PG_CP_THREADS 128
#endif
#ifndef PG_CP_MINBLOCKS
#define PG_CP_MINBLOCKS 4
#endif
constexpr int CP_THREADS = PG_CP_THREADS; // threads per block of fp_pass_poly
#ifndef PG_CP_STAGES
#define PG_CP_STAGES 4
#endif
#ifndef PG_CP_UNROLL
#define PG_CP_UNROLL 1 // rows per trip of the streaming loop (A/B builds)
#endif
constexpr int CP_UNROLL = PG_CP_UNROLL;
constexpr int CP_STAGES = PG_CP_STAGES; // rows of particle data in flight per warp (cp.async ring, 1.5 KB per stage)
constexpr int CP_STAGE_D2 = 96;  // double2 slots per stage: X, V, v pairs of the 32 lanes

__host__ __device__ inline size_t cp_smem_bytes(int threads)
{
    return (size_t)(threads / 32) * (CP_GS * CP_WG * sizeof(double) + (size_t)CP_STAGES * CP_STAGE_D2 * sizeof(double2)) +
           (size_t)threads * CP_NM * sizeof(double); // + the spare moment set of every lane (deterministic mode)
}

// Table row of interval m (any sign): m mod (N * CP_NSUB).  Its stencil centre is the Julia cell index k = (m+4)>>3
// (0-based cell z = k-1), its sub-interval s = (m+4)&7.  Conversely the row of (z, s) is 8z + s + 4.
__device__ __forceinline__ int cp_row_of(int z, int s, int Mmask) { return (CP_NSUB * z + s + CP_NSUB / 2) & Mmask; }

// G[row(z,s)*CP_GS + n] = sum_j CWS[s][j][n] * E[(z+j) mod N]; also clears the moment grid the previous mom2rho_kernel
// consumed.  gridDim.y = CP_NSUB: a block serves one sub-interval, so the table index is uniform (constant-bank operands).
// Skipped (like the solve) once the step has converged.
struct GPolyArgs {
    const double *E;
    double *G;   // [N*CP_NSUB][CP_GS]
    fx_t *Mg;    // [N*CP_NSUB][CP_NC]
    const Ctrl *ctrl;
    int N, k;
    int det;     // deterministic mode: Mg holds two words per moment
};

__global__ void __launch_bounds__(128) gpoly_kernel(GPolyArgs a)
{
    const int fk = a.ctrl->final_k;
    if (fk >= 0 && sweep_index(a.k, a.ctrl) > fk) return;
    const int N = a.N, Nmask = N - 1, Mmask = N * CP_NSUB - 1;
    const int z = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
    if (z < N) {
        double e[GAUSS_NW];
#pragma unroll
        for (int q = 0; q < GAUSS_NW; ++q) e[q] = a.E[(z + q - 6) & Nmask];
        double *g = a.G + (size_t)cp_row_of(z, s, Mmask) * CP_GS;
#pragma unroll
        for (int n = 0; n < CP_NC; ++n) {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < GAUSS_NW; ++q) acc = fma(PG_CWS[s][q][n], e[q], acc);
            g[n] = acc;
        }
        g[CP_NC] = 0.0;
    }
    const int nb = gridDim.x * gridDim.y, b = blockIdx.y * gridDim.x + blockIdx.x;
    const long long total = (long long)N * CP_NSUB * CP_NC * (a.det ? 2 : 1), per = (total + nb - 1) / nb;
    const long long lo = (long long)b * per, hi = min(total, lo + per);
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) a.Mg[i] = 0ULL;
}

// rho_fx[i] += fixed( sum_j sum_s sum_n CWS[s][j][n] * M[row(i-j, s)][n] ).  One block = CPM_CELLS output cells; the
// moments of the CPM_CELLS+12 source cells are converted to fp64 once in shared memory (cell stride odd: the lanes of a warp
// read different banks); warp s sums the contributions of sub-interval s (uniform table index), the 8 partial sums of a
// cell are added in a fixed order.  Mg is cleared later by gpoly_kernel (other blocks read the halo cells).
struct Mom2RhoArgs {
    const fx_t *Mg;
    fx_t *rho;
    const Ctrl *ctrl;
    double fx_scale, fx_inv;
    int N;
    const unsigned long long *flush_src; // NCCL path on several GPUs: this rank's flush counter goes to rho[N] and is summed with the grid
    int det;     // deterministic mode: Mg[row][n] = (low 32-bit sums, high-part sums) of 50-fractional-bit integers
};
constexpr int CPM_CELLS = 32;
constexpr int CP_DET_FRAC = 50;                      // fractional bits of a lane's integer moment sums
constexpr double CP_DET_MAGIC = 6.0;                 // 1.5 * 2^(52-50): |u^n| <= 1 keeps u^n + 6 inside [4, 8), where ulp = 2^-50
constexpr int CP_DET_MAXCNT = 2048;                  // particles per set before a forced flush: |sum| < 2^11 * 2^50 < 2^62
constexpr int CP_DETV_FRAC = 40;                     // the same for the diagnostics sums: v^2, v in (-2048, 2048)
constexpr double CP_DETV_MAGIC = 6144.0;             // 1.5 * 2^(52-40)
constexpr int CPM_LD = CP_NSUB * CP_NC + 1; // 89 doubles per source cell

constexpr size_t CPM_SMEM = (size_t)(CPM_CELLS + 12) * CPM_LD * sizeof(double); // dynamic shared memory of mom2rho_kernel

__global__ void __launch_bounds__(32 * CP_NSUB) mom2rho_kernel(Mom2RhoArgs a)
{
    extern __shared__ double Ms[]; // [(CPM_CELLS + 12) * CPM_LD]
    __shared__ double part[CP_NSUB][CPM_CELLS];
    if (a.ctrl->final_k >= 0) return; // converged: the moments belong to the next step's first solve
    const int N = a.N, Nmask = N - 1, Mmask = N * CP_NSUB - 1;
    const int i0 = blockIdx.x * CPM_CELLS;
    if (a.flush_src && blockIdx.x == 0 && threadIdx.x == 0) a.rho[N] = *a.flush_src;
    for (int t = threadIdx.x; t < (CPM_CELLS + 12) * CP_NSUB * CP_NC; t += blockDim.x) {
        const int c = t / (CP_NSUB * CP_NC), r = t - c * (CP_NSUB * CP_NC); // r = s*CP_NC + n: the rows of a cell are contiguous in Mg
        const int s = r / CP_NC, n = r - s * CP_NC;
        const size_t g = (size_t)cp_row_of((i0 + c - 6) & Nmask, s, Mmask) * CP_NC + n;
        if (a.det) // exact integers, put together in a fixed order: the same bits whatever order the lanes flushed in
            Ms[c * CPM_LD + r] = ((double)(long long)a.Mg[2 * g + 1] * 4294967296.0 + (double)a.Mg[2 * g]) * (1.0 / (double)(1LL << CP_DET_FRAC));
        else
            Ms[c * CPM_LD + r] = (double)(long long)a.Mg[g] * a.fx_inv;
    }
    __syncthreads();
    const int i = threadIdx.x & 31, s = threadIdx.x >> 5;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int q = 0; q < GAUSS_NW; ++q) {
        // cell i receives W_j from source cell i - j, j = q - 6; source slot = (i - j) + 6 = i + 12 - q
        const double *m = Ms + (i + 12 - q) * CPM_LD + s * CP_NC;
        double acc = 0.0;
#pragma unroll
        for (int n = 0; n < CP_NC; ++n) acc = fma(PG_CWS[s][q][n], m[n], acc);
        if (q & 1) s1 += acc; else s0 += acc;
    }
    part[s][i] = s0 + s1;
    __syncthreads();
    if (threadIdx.x < CPM_CELLS && i0 + i < N) {
        double r = 0.0; // the sub-intervals of a cell, added in a fixed order (8 intervals: the same tree as before)
#pragma unroll
        for (int s8 = 0; s8 < CP_NSUB; s8 += 8)
            r += ((part[s8][i] + part[s8 + 1][i]) + (part[s8 + 2][i] + part[s8 + 3][i])) + ((part[s8 + 4][i] + part[s8 + 5][i]) + (part[s8 + 6][i] + part[s8 + 7][i]));
        a.rho[i0 + i] += to_fx(r, a.fx_scale);
    }
}

// Interval m = round(y) (ties to even) of y = c*N*CP_NSUB, as an integer (the low mantissa word of y + 1.5*2^52) and as a
// double (the interval's centre).  No conversion-unit instructions.
__device__ __forceinline__ void cp_interval(double y, int &idx, double &centre)
{
    const double big = y + CP_MAGIC;
    idx = __double2loint(big);
    centre = big - CP_MAGIC;
}

// One lane's moment set for the interval it is currently in (n = 0 is an integer count).  DET: integer sums of the raw
// bit patterns of u^n + CP_DET_MAGIC.
template <bool DET>
struct CPSet {
    typename std::conditional<DET, unsigned long long, double>::type m[CP_NM];
    double centre;
    int cnt, idx;
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int n = 0; n < CP_NM; ++n) m[n] = 0;
        cnt = 0;
    }
};

// One moment value (fixed point, CP_DET_FRAC fractional bits) into the two-word deterministic grid.
__device__ __forceinline__ void cp_det_red(fx_t *p2, long long val)
{
    atomicAdd(p2, (fx_t)(val & 0xFFFFFFFFLL));
    atomicAdd(p2 + 1, (fx_t)(val >> 32));
}

// Deterministic mode: add a lane's set to the two-word moment grid (22 integer REDs) and clear it.  (The default mode flushes
// per warp: cp_flush_warp below.)
__device__ __forceinline__ void cp_flush(CPSet<true> &s, fx_t *Mg, int Mmask)
{
    if (s.cnt) {
        fx_t *p = Mg + (size_t)(s.idx & Mmask) * (2 * CP_NC);
        cp_det_red(p, (long long)s.cnt << CP_DET_FRAC);
        const unsigned long long off = (unsigned long long)s.cnt * (unsigned long long)__double_as_longlong(CP_DET_MAGIC);
#pragma unroll
        for (int n = 0; n < CP_NM; ++n) cp_det_red(p + 2 * (1 + n), (long long)((unsigned long long)s.m[n] - off));
        s.clear();
    }
}

// The spare set of the deterministic mode: count and interval in registers, the sums in the lane's shared-memory column.
struct CPSpare { int cnt, idx; };

__device__ __forceinline__ void cp_flush_spare(CPSpare &sp, unsigned long long *col, int stride, fx_t *Mg, int Mmask)
{
    if (sp.cnt) {
        fx_t *p = Mg + (size_t)(sp.idx & Mmask) * (2 * CP_NC);
        cp_det_red(p, (long long)sp.cnt << CP_DET_FRAC);
        const unsigned long long off = (unsigned long long)sp.cnt * (unsigned long long)__double_as_longlong(CP_DET_MAGIC);
#pragma unroll
        for (int n = 0; n < CP_NM; ++n) { cp_det_red(p + 2 * (1 + n), (long long)(col[n * stride] - off)); col[n * stride] = 0ULL; }
        sp.cnt = 0;
    }
}

// Deterministic mode: M[m][n] += u^n for the particle at y, m = round(y) whatever the lane did before; `col` / `sp` hold the
// other interval of an alternating lane.
__device__ __forceinline__ void cp_deposit1(double y, CPSet<true> &P, CPSpare &sp, unsigned long long *col, int stride, fx_t *Mg,
                                            int Mmask, unsigned int &nflush)
{
    int idx;
    double centre;
    cp_interval(y, idx, centre);
    const double u = y - centre;
    if (idx != P.idx || P.cnt >= CP_DET_MAXCNT) {
        if (idx == sp.idx && idx != P.idx) { // change over to the other interval: exchange the two sets
#pragma unroll
            for (int n = 0; n < CP_NM; ++n) { const unsigned long long m = col[n * stride]; col[n * stride] = P.m[n]; P.m[n] = m; }
            const int c = sp.cnt; sp.cnt = P.cnt; P.cnt = c;
            sp.idx = P.idx; P.idx = idx;
            ++nflush;
            if (P.cnt >= CP_DET_MAXCNT) cp_flush(P, Mg, Mmask);
        } else if (idx == P.idx) {           // full set
            cp_flush(P, Mg, Mmask);
        } else {                             // a third interval: retire the spare, park the current set
            nflush += sp.cnt > 0;
            cp_flush_spare(sp, col, stride, Mg, Mmask);
#pragma unroll
            for (int n = 0; n < CP_NM; ++n) { col[n * stride] = P.m[n]; P.m[n] = 0; }
            sp.cnt = P.cnt; sp.idx = P.idx;
            P.cnt = 0; P.idx = idx;
        }
    }
    P.cnt++;
    const double u2 = u * u;
    double pa = u, pb = u2;
#pragma unroll
    for (int n = 0; n < CP_NM; n += 2) {
        P.m[n] += (unsigned long long)__double_as_longlong(pa + CP_DET_MAGIC);
        P.m[n + 1] += (unsigned long long)__double_as_longlong(pb + CP_DET_MAGIC);
        if (n + 2 < CP_NM) { pa *= u2; pb *= u2; }
    }
}

// ---- default (not deterministic) deposit: the WARP shares one interval --------------------------------------------------
// Every lane sums u^n of its particles relative to the warp's current interval (centre, idx: warp-uniform) in registers.  When
// most of a row lies more than CP_UMOVE = 3/4 of an interval from the centre the stream has moved on: the 32 lane sets are added up through
// shared memory (lane n sums moment n) and go to the moment grid with ONE set of 11 REDs, and the warp re-centres on the
// interval of the first such particle.  With every lane flushing for itself (the form this replaces) the same event cost
// 32 x 11 REDs to the same 11 addresses -- 0.35 ms of a 5 ms step at 2^28 particles (measured by issuing them twice).
// The 3/4 is hysteresis: a bin that straddles an interval edge does not make the warp alternate.  A particle
// more than CP_UMAX = 1 interval from the centre (outside the range the polynomials are fitted on -- disordered input)
// is deposited on its own (11 REDs).
__device__ __noinline__ void cp_deposit_single(double y, fx_t *Mg, double fx_scale, int Mmask)
{
    int idx;
    double centre;
    cp_interval(y, idx, centre);
    const double u = y - centre, u2 = u * u;
    fx_t *p = Mg + (size_t)(idx & Mmask) * CP_NC;
    atomicAdd(p, to_fx(1.0, fx_scale));
    double pa = u, pb = u2;
#pragma unroll
    for (int n = 0; n < CP_NM; n += 2) {
        atomicAdd(p + 1 + n, to_fx(pa, fx_scale));
        atomicAdd(p + 2 + n, to_fx(pb, fx_scale));
        pa *= u2; pb *= u2;
    }
}

// colw: the warp's [CP_NM][32] slice of the lane-column area (row stride cstride doubles).
__device__ __forceinline__ void cp_flush_warp(CPSet<false> &s, double *colw, int lane, int cstride, fx_t *Mg, double fx_scale, int Mmask)
{
    const int total = __reduce_add_sync(0xffffffffu, s.cnt);
    if (total) {
#pragma unroll
        for (int n = 0; n < CP_NM; ++n) colw[n * cstride + lane] = s.m[n];
        __syncwarp();
        fx_t *p = Mg + (size_t)(s.idx & Mmask) * CP_NC;
        if (lane < CP_NM) {
            const double *r = colw + lane * cstride;
            double t0 = 0.0, t1 = 0.0;
#pragma unroll 8
            for (int i = 0; i < 32; i += 2) { t0 += r[(i + lane) & 31]; t1 += r[(i + 1 + lane) & 31]; } // rotated: no bank conflicts
            atomicAdd(p + 1 + lane, to_fx(t0 + t1, fx_scale));
        } else if (lane == CP_NM) {
            atomicAdd(p, to_fx((double)total, fx_scale));
        }
        __syncwarp();
        s.clear();
    }
}

// M += u^n, n = 1..CP_NM, and the count, for one particle of the lane.
__device__ __forceinline__ void cp_accumulate(CPSet<false> &P, double u)
{
    P.cnt++;
    const double u2 = u * u;
    double pa = u, pb = u2;
#pragma unroll
    for (int n = 0; n < CP_NM; n += 2) {
        P.m[n] += pa; P.m[n + 1] += pb;
        if (n + 2 < CP_NM) { pa *= u2; pb *= u2; }
    }
}

__device__ __forceinline__ void cp_deposit_row(const double (&y)[2], CPSet<false> &P, double *colw, int lane, int cstride, fx_t *Mg,
                                               double fx_scale, int Mmask, unsigned int &nflush)
{
    double u[2] = {y[0] - P.centre, y[1] - P.centre};
    const int far = (fabs(u[0]) <= CP_UMOVE ? 0 : 1) + (fabs(u[1]) <= CP_UMOVE ? 0 : 1);
    const int nfar = __reduce_add_sync(0xffffffffu, far); // one REDUX; most rows of a sorted stream: 0
    if (nfar == 0) {
        cp_accumulate(P, u[0]);
        cp_accumulate(P, u[1]);
        return;
    }
    if (nfar > 32) { // most of the row has left the interval (or there is none yet)
        nflush += P.cnt > 0;
        cp_flush_warp(P, colw, lane, cstride, Mg, fx_scale, Mmask);
        const unsigned int f0 = __ballot_sync(0xffffffffu, !(fabs(u[0]) <= CP_UMOVE)), f1 = __ballot_sync(0xffffffffu, !(fabs(u[1]) <= CP_UMOVE));
        const double yr = __shfl_sync(0xffffffffu, f0 ? y[0] : y[1], __ffs(f0 ? f0 : f1) - 1);
        cp_interval(yr, P.idx, P.centre);
        u[0] = y[0] - P.centre; u[1] = y[1] - P.centre;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (fabs(u[q]) <= CP_UMAX) {
            cp_accumulate(P, u[q]);
        } else {
            ++nflush;
            cp_deposit_single(y[q], Mg, fx_scale, Mmask);
        }
    }
}

// sum_n u^n g[n] from one row (128-bit loads of the pairs (c_2m, c_2m+1)): even and odd halves as two independent
// Horner chains in u^2.
__device__ __forceinline__ double cp_horner(const double *g, double u)
{
    static_assert(CP_NC % 2 == 1 && CP_NC >= 5, "cp_horner is written for an even degree");
    constexpr int H = (CP_NC - 1) / 2;
    const double2 *g2 = reinterpret_cast<const double2 *>(g);
    const double u2 = u * u;
    double2 c = g2[H];
    double ge = c.x; // top coefficient (even)
    c = g2[H - 1];
    ge = fma(ge, u2, c.x);
    double go = c.y; // top odd coefficient
#pragma unroll
    for (int m = H - 2; m >= 0; --m) {
        c = g2[m];
        ge = fma(ge, u2, c.x);
        go = fma(go, u2, c.y);
    }
    return fma(go, u, ge);
}

// Rare path: the particle's interval is outside the warp's gather window (same operation order as cp_horner).
__device__ __noinline__ double cp_slow_gather(const double *G, int idx, double u, int Mmask)
{
    return cp_horner(G + (size_t)(idx & Mmask) * CP_GS, u);
}

// Diagnostics sums of the final pass: sum(v.^2), sum(v) (GaussianFixedPoint.jl:10).  DET: integer sums of the raw patterns of
// v^2 + 6144 and v + 6144 (CP_DETV_FRAC fractional bits), order-free like the moments.
template <bool DET>
struct CPVSum {
    double s2 = 0.0, s1 = 0.0;
    unsigned long long i2 = 0ULL, i1 = 0ULL;
    long long n = 0;
    __device__ __forceinline__ void add(double v)
    {
        if (DET) {
            i2 += (unsigned long long)__double_as_longlong(v * v + CP_DETV_MAGIC);
            i1 += (unsigned long long)__double_as_longlong(v + CP_DETV_MAGIC);
            ++n;
        } else {
            s2 = fma(v, v, s2); s1 += v;
        }
    }
};

// Block-wide sum of 64-bit integers; result valid in thread 0.  scratch: >= 32 x 8 bytes.
__device__ __forceinline__ long long block_sum_ll(long long v, double *scratch)
{
    long long *sc = reinterpret_cast<long long *>(scratch);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sc[wid] = v;
    __syncthreads();
    long long r = 0;
    if (wid == 0) {
        r = lane < nw ? sc[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    return r;
}

// ---- fused re-sort (see the header) ----------------------------------------------------------------------------------------
// Bin of the particle that ends the step at x = xn (not yet wrapped) with velocity vb: position xn + vb*dt/2 rounded to the nearest
// 1/2^sublg of a cell, offset by half a cell (a cell's bins surround its stencil centre, like sort_key_of), wrapped; sign of vb.
__device__ __forceinline__ int cp_fs_key(const FPArgs &a, double xn, double vb)
{
    const double big = fma(vb, a.fs_hs, xn * a.fs_scale) + a.fs_magic;
    const int I = __double2loint(big);
    const int c = (I >> a.fs_sublg) & (a.N - 1);
    return (((c << 1) | (vb >= 0.0 ? 1 : 0)) << a.fs_sublg) | (I & ((1 << a.fs_sublg) - 1));
}

// Counting costs a pass ~25 % more instructions, so only the passes that may precede the final one count: pass k does if
// k + 1 >= the smaller sweep count of the last two steps (Ctrl.fs_pred, set by step_end_kernel; 0 after a reset: every pass
// counts).  A step that ends earlier than that finds no counts and writes unpermuted (one re-sort is skipped, nothing else).
__device__ __forceinline__ bool cp_fs_counts(int k, const Ctrl *c) { return k >= 1 && k + 1 >= c->fs_pred; }

// Lanes holding equal keys next to each other form a run: its first lane and its length (the whole warp calls this).
__device__ __forceinline__ void cp_fs_runs(int key, int lane, int &head, int &len)
{
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned int H = __ballot_sync(0xffffffffu, lane == 0 || key != prev);
    head = 31 - __clz(H & (0xffffffffu >> (31 - lane)));
    const unsigned int above = H & ~((2u << head) - 1u);
    len = (above ? __ffs(above) - 1 : 32) - head;
}

// Count the two particles of every lane of a row.  In sorted order a lane's pair shares its bin (the warp votes): then runs are
// taken over the lanes and count double; otherwise the even and the odd particles of the row form runs of their own.
__device__ __forceinline__ void cp_fs_count(const FPArgs &a, const double (&xn)[2], const double (&vb)[2], int lane)
{
    const int k0 = cp_fs_key(a, xn[0], vb[0]), k1 = cp_fs_key(a, xn[1], vb[1]);
    const bool pairs = __all_sync(0xffffffffu, k0 == k1);
    int head, len;
    cp_fs_runs(k0, lane, head, len);
    if (lane == head) atomicAdd(&a.fs_hist[k0], (unsigned int)(pairs ? 2 * len : len));
    if (!pairs) {
        cp_fs_runs(k1, lane, head, len);
        if (lane == head) atomicAdd(&a.fs_hist[k1], (unsigned int)len);
    }
}

// Slot reservation of a row, in two halves so that the atomics are in flight while the row's gather is evaluated: reserve()
// right after the row is read, slots() just before the stores.
struct CPFsRes {
    unsigned int base[2];
    int head[2];
    bool pairs;
};
__device__ __forceinline__ void cp_fs_reserve(const FPArgs &a, const double (&xn)[2], const double (&vb)[2], int lane, CPFsRes &r)
{
    const int k0 = cp_fs_key(a, xn[0], vb[0]), k1 = cp_fs_key(a, xn[1], vb[1]);
    r.pairs = __all_sync(0xffffffffu, k0 == k1);
    int len;
    cp_fs_runs(k0, lane, r.head[0], len);
    r.base[0] = r.base[1] = 0u;
    if (lane == r.head[0]) r.base[0] = atomicAdd(&a.fs_cursor[k0], (unsigned int)(r.pairs ? 2 * len : len));
    r.head[1] = r.head[0];
    if (!r.pairs) {
        cp_fs_runs(k1, lane, r.head[1], len);
        if (lane == r.head[1]) r.base[1] = atomicAdd(&a.fs_cursor[k1], (unsigned int)len);
    }
}
__device__ __forceinline__ void cp_fs_slots(const CPFsRes &r, int lane, long long (&dst)[2])
{
    const unsigned int b0 = __shfl_sync(0xffffffffu, r.base[0], r.head[0]);
    if (r.pairs) {
        dst[0] = (long long)b0 + 2 * (lane - r.head[0]);
        dst[1] = dst[0] + 1;
    } else {
        const unsigned int b1 = __shfl_sync(0xffffffffu, r.base[1], r.head[1]);
        dst[0] = (long long)b0 + (lane - r.head[0]);
        dst[1] = (long long)b1 + (lane - r.head[1]);
    }
}

// Between the solve of sweep k and its pass: sweep k final -> exclusive scan of the bin counts of pass k-1 into the slot cursors;
// otherwise the counts are stale (pass k counts afresh).  Leaves the counts zero either way.  One block per chunk of 16384 bins
// (16 per thread: 4 x 128-bit loads); a block publishes its total as (epoch, total) in one 64-bit word and adds up the words
// of the blocks below it -- at most FS_MAXBLOCKS blocks, all resident, lower blocks never wait for higher ones.  The epoch is
// unique per (step, sweep) since the last reset (which clears the words).
struct FsScanArgs {
    unsigned int *hist, *cursor;
    unsigned long long *sync; // [FS_MAXBLOCKS]
    Ctrl *ctrl;
    int nbins, k;
};
constexpr int FS_CHUNK = 1024 * 16;
constexpr int FS_MAXBLOCKS = 32;

__global__ void __launch_bounds__(1024) cp_fs_scan_kernel(FsScanArgs a)
{
    __shared__ unsigned int wsum[32];
    __shared__ unsigned int carry_s;
    if (a.k < 0 && blockIdx.x == 0 && threadIdx.x == 0) a.ctrl->loop_fs_sweeps += 1ULL; // launch accounting of the device-driven loop
    const int fk = a.ctrl->final_k, k = sweep_index(a.k, a.ctrl);
    if (k < 2 || (fk >= 0 && k > fk)) return; // pass 0 does not count: at sweep 1 the counts are still zero
    if (!cp_fs_counts(k - 1, a.ctrl)) return; // neither did pass k-1 (the step was expected to take more sweeps)
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5, b = blockIdx.x;
    const int nbins = a.nbins, base = b * FS_CHUNK; // nbins: a multiple of 4
    uint4 c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = base + 16 * t + 4 * i;
        if (idx < nbins) {
            if (fk == k) c[i] = *reinterpret_cast<const uint4 *>(a.hist + idx);
            *reinterpret_cast<uint4 *>(a.hist + idx) = make_uint4(0u, 0u, 0u, 0u);
        } else {
            c[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    if (fk != k) return;
    unsigned int s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { // exclusive scan of the thread's 16 counts, in place
        unsigned int v;
        v = c[i].x; c[i].x = s; s += v;
        v = c[i].y; c[i].y = s; s += v;
        v = c[i].z; c[i].z = s; s += v;
        v = c[i].w; c[i].w = s; s += v;
    }
    unsigned int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const unsigned int w = wsum[lane];
        unsigned int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += u;
        }
        wsum[lane] = wi - w;
        const unsigned long long epoch = (unsigned long long)((unsigned int)(a.ctrl->step * 64 + k) + 1u) << 32;
        volatile unsigned long long *sync = a.sync;
        if (lane == 31) sync[b] = epoch | wi; // this block's total
        unsigned int below = 0;
        if (lane < b) {
            unsigned long long w64;
            do { w64 = sync[lane]; } while ((w64 >> 32) != (epoch >> 32));
            below = (unsigned int)w64;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        if (lane == 0) carry_s = below;
    }
    __syncthreads();
    const unsigned int off = carry_s + wsum[wid] + inc - s;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = base + 16 * t + 4 * i;
        if (idx < nbins) *reinterpret_cast<uint4 *>(a.cursor + idx) = make_uint4(off + c[i].x, off + c[i].y, off + c[i].z, off + c[i].w);
    }
}

// The last P % 64 particles of a shard (no full row): one thread each, global-memory gather polynomial and direct
// moment REDs.  Same arithmetic as the streaming loop below.
template <bool FIRST, bool DET>
__device__ __forceinline__ void cp_tail_particle(const FPArgs &a, long long j, bool final, bool v0_is_V, CPVSum<DET> &vs)
{
    const int Mmask = a.N * CP_NSUB - 1;
    const double dNs = a.dN, hdt = a.dt / 2;
    double Xj = a.X[j], Vj = a.V[j], vj = v0_is_V ? Vj : a.v[j];
    double xj = Xj + (vj + Vj) * hdt;
    int idx;
    double centre;
    if (!FIRST) {
        const bool fs = a.fs_hist != nullptr;
        long long dst = j; // fused re-sort: slot of this particle's end-of-step state (counted by the previous pass from the same operands)
        const int k = sweep_index(a.k, a.ctrl);
        if (fs && final && k >= 2 && cp_fs_counts(k - 1, a.ctrl)) dst = (long long)atomicAdd(&a.fs_cursor[cp_fs_key(a, xj, vj)], 1u);
        const double y = (xj + Xj) * dNs;
        cp_interval(y, idx, centre);
        vj = Vj + cp_slow_gather(a.G, idx, y - centre, Mmask) * a.dt;
        if (final && fs) { a.fs_vout[dst] = vj; a.fs_pid_out[dst] = a.fs_pid_in[j]; }
        else a.v[j] = vj;
        if (final) {
            const double xw = jl_mod1(xj);
            a.xout[dst] = xw;
            vs.add(vj);
            Xj = xw; Vj = vj;
        }
        xj = Xj + (vj + Vj) * hdt;
        if (fs && !final && cp_fs_counts(k, a.ctrl)) atomicAdd(&a.fs_hist[cp_fs_key(a, xj, vj)], 1u);
    }
    const double y = (xj + Xj) * dNs;
    if constexpr (!DET) {
        cp_deposit_single(y, a.Mg, a.fx_scale, Mmask);
    } else { // the streaming loop's own arithmetic (same powers, same rounding) for a set of one particle
        CPSet<true> one;
        one.clear(); one.centre = 1e300; one.idx = 0x40000000; // not an interval any particle can be in
        unsigned int nf = 0;
        CPSpare nosp; nosp.cnt = 0; nosp.idx = 0x40000001;
        unsigned long long dummy[CP_NM];
        cp_deposit1(y, one, nosp, dummy, 1, a.Mg, Mmask, nf);
        cp_flush(one, a.Mg, Mmask);
    }
}

// Which of the run-time properties of a pass the streaming loop of fp_pass_poly is compiled for (-1: read at run time).  The kernel
// nodes of the device-driven loop are the same for every sweep, so the pass learns on the device whether it is the first sweep
// (v = V: no work-buffer stream), a middle one or the final one, and whether the step re-sorts; the loop body is instantiated for the
// three plain cases with these as constants (fewer uniform branches, selects and live values in a loop that is bound by
// instruction issue) and once generically for everything else.
template <int F, int V, int S>
struct CPSpec { static constexpr int final_ = F, v0 = V, fs = S; };

// Pass k of a step (same contract as fp_pass_sorted / fp_pass_atomic; FPArgs.G / FPArgs.Mg carry the polynomial
// tables, FPArgs.dN = N*CP_NSUB/2 so that y = (x+X)*dN).  Row = 64 consecutive particles; lane l owns particles 2l and 2l+1.
template <bool FIRST, bool DET>
__global__ void __launch_bounds__(CP_THREADS, PG_CP_MINBLOCKS) fp_pass_poly(FPArgs a)
{
    extern __shared__ double smem[];
    __shared__ double scratch[32];
    const int fk = a.ctrl->final_k, k = FIRST ? 0 : sweep_index(a.k, a.ctrl);
    if (!FIRST && fk >= 0 && k > fk) return;
    const bool final_rt = !FIRST && fk == k;
    const bool v0_rt = FIRST || k == 1; // sweep 1 starts from v = V: the work buffer is stale until pass 1 writes it
    const bool fs_rt = !FIRST && a.fs_hist != nullptr; // re-sort fused into this step's passes
    const bool fs_scatter_rt = fs_rt && final_rt && k >= 2 && cp_fs_counts(k - 1, a.ctrl), fs_count_rt = fs_rt && !final_rt && cp_fs_counts(k, a.ctrl);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double *Gw = smem + warp * (CP_GS * CP_WG); // [CP_WG][CP_GS]
    const int Mmask = a.N * CP_NSUB - 1;
    const double dNs = a.dN, hdt = a.dt / 2; // ((v+V)/2)*dt == (v+V)*(dt/2) bit for bit (exact power-of-two scalings)
    // full rows only; every warp streams one contiguous range [r0, r1)
    const int rows = (int)(a.P >> 6);
    const int nw = gridDim.x * wpb, gw = blockIdx.x * wpb + warp;
    const int rpw = (rows + nw - 1) / nw;
    const int r0 = min(rows, gw * rpw), r1 = min(rows, r0 + rpw);
    const double2 *X2 = reinterpret_cast<const double2 *>(a.X), *V2 = reinterpret_cast<const double2 *>(a.V);
    double2 *v2 = reinterpret_cast<double2 *>(a.v), *xo2 = reinterpret_cast<double2 *>(a.xout);
    CPSet<DET> A;
    A.clear(); A.idx = 0x40000000; A.centre = 1e300; // no interval yet: the first particle opens one
    CPSpare spare; spare.cnt = 0; spare.idx = 0x40000001;
    unsigned long long *col = reinterpret_cast<unsigned long long *>(smem + wpb * (CP_GS * CP_WG + 2 * CP_STAGES * CP_STAGE_D2)) + threadIdx.x;
    const int cstride = blockDim.x;
    if (DET) {
#pragma unroll
        for (int n = 0; n < CP_NM; ++n) col[n * cstride] = 0ULL;
    }
    int gb = 0x40000000; // interval of window slot 0; the first row always restages (see `staged`)
    bool staged = false;
    CPVSum<DET> vs;
    unsigned int nflush = 0;
    // particle rows arrive through a per-warp cp.async ring: no registers are held while a row is in flight
    double2 *ring = reinterpret_cast<double2 *>(smem + wpb * (CP_GS * CP_WG)) + warp * (CP_STAGES * CP_STAGE_D2) + lane;
    auto stream = [&](auto spec) {
    using Spec = decltype(spec);
    const bool final = Spec::final_ < 0 ? final_rt : Spec::final_ != 0;
    const bool v0_is_V = Spec::v0 < 0 ? v0_rt : Spec::v0 != 0;
    const bool fs = Spec::fs < 0 ? fs_rt : false;
    const bool fs_final = fs && final, fs_scatter = Spec::fs < 0 ? fs_scatter_rt : false, fs_count = Spec::fs < 0 ? fs_count_rt : false;
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
#pragma unroll CP_UNROLL
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
            uint2 ids = make_uint2(0u, 0u);
            CPFsRes res;
            if (fs_final) { // fused re-sort: reserve this row's slots early, the atomics return while the gather is evaluated
#ifndef PG_EXP_NOPID
                ids = __ldcs(reinterpret_cast<const uint2 *>(a.fs_pid_in) + j2);
#endif
                if (fs_scatter) cp_fs_reserve(a, xj, vj, lane, res);
            }
            int idx[2];
            double u[2];
            unsigned int slot[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const double y = (xj[q] + Xj[q]) * dNs;
                double centre;
                cp_interval(y, idx[q], centre);
                u[q] = y - centre;
                slot[q] = (unsigned int)(idx[q] - gb) & (unsigned int)Mmask;
            }
            if (!__all_sync(0xffffffffu, staged && slot[0] < CP_WG && slot[1] < CP_WG)) {
                // recentre the window two intervals below the smallest interval of this row and restage it
                int cm = min(idx[0], idx[1]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cm = min(cm, __shfl_xor_sync(0xffffffffu, cm, o));
                gb = cm - 2;
                staged = true;
                __syncwarp();
                for (int i = lane; i < CP_GS * CP_WG; i += 32) {
                    const int s = i / CP_GS, n = i - s * CP_GS;
                    Gw[i] = a.G[(size_t)((gb + s) & Mmask) * CP_GS + n];
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 2; ++q) slot[q] = (unsigned int)(idx[q] - gb) & (unsigned int)Mmask;
            }
            double g[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) g[q] = cp_horner(Gw + (slot[q] < CP_WG ? slot[q] : 0u) * CP_GS, u[q]);
            if (slot[0] >= CP_WG || slot[1] >= CP_WG) { // rare: an interval more than CP_WG above the row's smallest
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (slot[q] >= CP_WG) g[q] = cp_slow_gather(a.G, idx[q], u[q], Mmask);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) vj[q] = Vj[q] + g[q] * a.dt; // v[j]=V[j]+sum(...)*dt
            if (!fs_final) __stcs(v2 + j2, make_double2(vj[0], vj[1]));
            if (final) {
                // end of step: x.=mod.(x,1), diagnostics sums -- and the first pass of the NEXT step fused in
                // (X.=x; V.=v; x = X + (V+V)/2*dt; deposit at (x+X)/2)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    Xj[q] = jl_mod1(xj[q]);
                    vs.add(vj[q]);
                    Vj[q] = vj[q];
                }
                if (fs) { // to the slots of the re-sorted arrays (runs of a row are contiguous)
                    long long dst[2] = {2 * j2, 2 * j2 + 1};
                    if (fs_scatter) cp_fs_slots(res, lane, dst);
#ifdef PG_EXP_NOSCATTER // measurement builds only: slots reserved, results written in stream order -- what the scattered stores cost
                    if (dst[0] != -1) { dst[0] = 2 * j2; dst[1] = 2 * j2 + 1; }
#endif
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        a.xout[dst[q]] = Xj[q];
                        a.fs_vout[dst[q]] = vj[q];
#ifndef PG_EXP_NOPID // measurement builds only (particles come back in sorted order): what carrying the ids costs
                        a.fs_pid_out[dst[q]] = q ? ids.y : ids.x;
#endif
                    }
                } else {
                    __stcs(xo2 + j2, make_double2(Xj[0], Xj[1]));
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) xj[q] = Xj[q] + (vj[q] + Vj[q]) * hdt;
            if (fs_count) cp_fs_count(a, xj, vj, lane); // the bins this row is written to if the next sweep is the final one
        }
        if constexpr (DET) {
#pragma unroll
            for (int q = 0; q < 2; ++q) cp_deposit1((xj[q] + Xj[q]) * dNs, A, spare, col, cstride, a.Mg, Mmask, nflush);
        } else {
            const double yd[2] = {(xj[0] + Xj[0]) * dNs, (xj[1] + Xj[1]) * dNs};
            cp_deposit_row(yd, A, reinterpret_cast<double *>(col) - lane, lane, cstride, a.Mg, a.fx_scale, Mmask, nflush);
        }
    }
    }; // stream
    if constexpr (FIRST) {
        stream(CPSpec<0, 1, 0>{});
    } else {
        if (fs_rt || (final_rt && v0_rt)) stream(CPSpec<-1, -1, -1>{}); // a re-sorting step; a step that ends at sweep 1
        else if (final_rt) stream(CPSpec<1, 0, 0>{});
        else if (v0_rt) stream(CPSpec<0, 1, 0>{});
        else stream(CPSpec<0, 0, 0>{});
    }
    const bool final = final_rt, v0_is_V = v0_rt;
    cp_async_wait<0>();
    if constexpr (DET) cp_flush(A, a.Mg, Mmask);
    else cp_flush_warp(A, reinterpret_cast<double *>(col) - lane, lane, cstride, a.Mg, a.fx_scale, Mmask);
    if (DET) cp_flush_spare(spare, col, cstride, a.Mg, Mmask);
    // ragged tail of the shard: fewer than 64 particles, first warp of the last block
    if (blockIdx.x == gridDim.x - 1 && warp == 0) {
        const long long j = ((long long)rows << 6) + lane;
        if (j < a.P) cp_tail_particle<FIRST, DET>(a, j, final, v0_is_V, vs);
        if (j + 32 < a.P) cp_tail_particle<FIRST, DET>(a, j + 32, final, v0_is_V, vs);
    }
    if (final) {
        if (DET) { // four integers per block: low 32 bits and high part of the two fixed-point sums (step_end_kernel adds them up)
            const unsigned long long off = (unsigned long long)vs.n * (unsigned long long)__double_as_longlong(CP_DETV_MAGIC);
            const long long v2 = (long long)(vs.i2 - off), v1 = (long long)(vs.i1 - off);
            const long long q0 = block_sum_ll(v2 & 0xFFFFFFFFLL, scratch), q1 = block_sum_ll(v2 >> 32, scratch);
            const long long q2 = block_sum_ll(v1 & 0xFFFFFFFFLL, scratch), q3 = block_sum_ll(v1 >> 32, scratch);
            if (threadIdx.x == 0) {
                long long *pp = reinterpret_cast<long long *>(a.partials) + 4 * blockIdx.x;
                pp[0] = q0; pp[1] = q1; pp[2] = q2; pp[3] = q3;
            }
        } else {
            const double sv2 = block_sum(vs.s2, scratch), sv = block_sum(vs.s1, scratch);
            if (threadIdx.x == 0) { a.partials[2 * blockIdx.x] = sv2; a.partials[2 * blockIdx.x + 1] = sv; }
        }
    }
    if (nflush && a.slow_count) atomicAdd(a.slow_count, (unsigned long long)nflush);
}

} // namespace pg
