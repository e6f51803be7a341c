// pg_kernels_1d.cuh -- particle passes of the 1D1V schemes.
//   NGP leapfrog        src/NGPFourier.jl:5-6
//   Gaussian leapfrog   src/Gaussian.jl:9-10
//   Gaussian fixed point src/GaussianFixedPoint.jl:7-10, src/GaussianFixedPointQuiet.jl:8-11
// Particle arrays are fp64 structure-of-arrays, read and written once per pass with coalesced
// streaming accesses.  Every pass keeps a private copy of E (gather) and of the deposit grid in
// shared memory and flushes the grid once per block.
#pragma once
#include "pg_common.cuh"
#include "pg_fft.cuh"
#include "pg_gauss.cuh"
#include "pg_tma.cuh"

namespace pg {

// ------------------------------------------------------------------------------------------
// deposit primitives on a block-private shared-memory grid
// ------------------------------------------------------------------------------------------
// All Gaussian deposits accumulate in 64-bit FIXED POINT: value = round(weight * 2^frac) with
// frac = 62 - ceil(log2(P+1)) so that even all P particles in one cell cannot overflow (picgolf.cu).
// Integer addition is
// associative, so the grid is bit-identical whatever the order of the atomics, the particle order or the
// number of GPUs -- in particular two sweeps with identical particle positions give identical rho, which
// is what lets the fixed point stop after 2 sweeps while the field is still round-off (as the sequential
// reference does).  In cell-sorted mode only whole chunk sums are quantised (relative error ~1e-14 at
// P = 2^28, N = 4096, below the ~1e-13 rounding error of a sequential fp64 running sum of 2^16 terms).
// rho = fixed * 2^-frac * w is formed by the solve.
__device__ __forceinline__ void gauss_deposit_atomic(fx_t *rs, int ibase, const double (&W)[GAUSS_NW], double fx_scale,
                                                     int Nmask)
{
    // r[k[1]] += k[2]*w      src/GaussianFixedPoint.jl:6   (the factor w is applied by the solve)
#pragma unroll
    for (int k = 0; k < GAUSS_NW; ++k) atomicAdd(&rs[(ibase + k - 1) & Nmask], to_fx(W[k], fx_scale)); // CAS loop in shared memory; see smem_add64
}

__device__ __forceinline__ void flush_grid(const fx_t *rs, fx_t *rho, int N)
{
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        fx_t r = rs[n];
        if (r) atomicAdd(&rho[n], r);
    }
}

// ------------------------------------------------------------------------------------------
// Gaussian fixed point.  Pass k of a step (k = 0..max_sweeps):
//   k = 0 (FIRST): x_0 = X + ((V+V)/2)*dt; deposit at (x_0+X)/2.
//   k >= 1: gather E_k at (x_{k-1}+X)/2 -> v_k = V + g*dt.  If the solve that produced E_k declared
//           the step finished (ctrl->final_k == k): x = mod(x_{k-1},1), accumulate sum v^2, sum v -- and, fused in,
//           the k = 0 pass of the NEXT step (deposit at x + v*dt/2), so every later step starts at its first solve.
//           Otherwise x_k = X + ((v_k+V)/2)*dt and deposit at (x_k+X)/2 for solve k+1.
//   Passes with k > final_k are predicated no-ops (the host never reads the flag mid-step).
// Note x lags v by one sweep on exit, exactly as in the reference (GaussianFixedPoint.jl:8-9).
// ------------------------------------------------------------------------------------------
struct FPArgs {
    const double *X, *V; // step-start state (X.=x; V.=v)
    double *v;           // working velocity v_k (in place)
    double *xout;        // wrapped end-of-step position
    const double *E;     // field of solve k
    fx_t *rho;           // global fixed-point deposit grid (zeroed by the solve that consumed it)
    fx_t *rho_next;      // grid of the NEXT step's first solve: target of the deposit fused into the final pass
    double *partials;    // [2*gridDim.x] per-block (sum v^2, sum v)
    Ctrl *ctrl;
    unsigned long long *slow_count; // sorted mode: particles that fell outside their warp's window
    long long P;
    double dt, fx_scale; // fx_scale = 2^frac
    int N, k;
    int K;               // sorted mode: batches of 32 particles per warp chunk
    const double *G;     // polynomial mode: per-cell gather polynomials [N][18] (pg_kernels_poly.cuh)
    fx_t *Mg;            // polynomial mode: fixed-point moment grid [N][17]
    double dN;           // (double)N
    // polynomial mode, re-sort fused into the passes of this step (pg_kernels_poly.cuh); fs_hist == NULL: off
    unsigned int *fs_hist, *fs_cursor; // bin counts written by the passes k >= 1, slot cursors scanned from them before a final pass
    double *fs_vout;                   // the final pass writes v to its slot here (the work buffer is still being read in place)
    const unsigned int *fs_pid_in;     // original index of every particle (picgolf_get_particles un-sorts with it)
    unsigned int *fs_pid_out;
    double fs_scale, fs_hs, fs_magic;  // N * 2^sublg, dt/2 * fs_scale, 1.5 * 2^52 + 2^(sublg-1)
    int fs_sublg;
};

template <bool FIRST>
__global__ void __launch_bounds__(PG_THREADS) fp_pass_atomic(FPArgs a)
{
    extern __shared__ double smem[];
    double *Es = smem, *scratch = smem + 2 * a.N;
    fx_t *rs = reinterpret_cast<fx_t *>(smem + a.N);
    const int fk = a.ctrl->final_k, k = FIRST ? 0 : sweep_index(a.k, a.ctrl);
    if (!FIRST && fk >= 0 && k > fk) return;
    const bool final = !FIRST && fk == k;
    const bool v0_is_V = FIRST || k == 1; // sweep 1 starts from v = V (`V.=v`): the work buffer is stale until pass 1 writes it
    const int N = a.N, Nmask = N - 1;
    const double dN = (double)N, dt = a.dt;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        if (!FIRST) Es[n] = a.E[n];
        rs[n] = 0ULL;
    }
    __syncthreads();
    double sv2 = 0.0, sv = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.P; j += stride) {
        const double Xj = ld_stream(a.X + j), Vj = ld_stream(a.V + j);
        double vj = v0_is_V ? Vj : ld_stream(a.v + j);
        double xj = Xj + (vj + Vj) / 2 * dt; // x.=X.+(v.+V)/2*dt
        int ibase;
        double W[GAUSS_NW];
        if (!FIRST) {
            gauss_weights((xj + Xj) / 2, dN, ibase, W);
            double g = gauss_gather(Es, ibase, W, Nmask);
            vj = Vj + g * dt; // v[j]=V[j]+sum(...)*dt
            st_stream(a.v + j, vj);
            if (final) {
                const double xw = jl_mod1(xj); // x.=mod.(x,1)
                st_stream(a.xout + j, xw);
                sv2 = fma(vj, vj, sv2);
                sv += vj;
                // fused first pass of the NEXT step (X.=x; V.=v; x = X + (V+V)/2*dt; deposit at (x+X)/2)
                xj = xw + (vj + vj) / 2 * dt;
                gauss_weights((xj + xw) / 2, dN, ibase, W);
                gauss_deposit_atomic(rs, ibase, W, a.fx_scale, Nmask);
                continue;
            }
            xj = Xj + (vj + Vj) / 2 * dt;
        }
        gauss_weights((xj + Xj) / 2, dN, ibase, W);
        gauss_deposit_atomic(rs, ibase, W, a.fx_scale, Nmask);
    }
    __syncthreads();
    if (final) {
        sv2 = block_sum(sv2, scratch);
        sv = block_sum(sv, scratch);
        if (threadIdx.x == 0) { a.partials[2 * blockIdx.x] = sv2; a.partials[2 * blockIdx.x + 1] = sv; }
    }
    flush_grid(rs, final ? a.rho_next : a.rho, N);
}

// ------------------------------------------------------------------------------------------
// Sorted variant of the fixed-point pass.  Particles are kept in cell order (pg_sort.cuh), so the 32*K
// consecutive particles a warp processes ("chunk") touch only a few neighbouring cells.  Each warp owns
// a WIN_ROWS-cell window of the grid in shared memory with one private column per lane:
//     acc[row][lane]  (row stride 32 doubles: a lane always hits its own bank whatever row it uses;
//                      the row sums of the flush read skewed columns, also conflict-free)
// so the 13 deposits of a particle are plain LDS/DADD/STS -- no atomics, no warp divergence; the
// SORTED_NP particles a lane evaluates together are summed in registers first when they share a window row.  At the
// end of the chunk lane r sums row r and issues ONE global fp64 RED per window cell.  The 32-cell slice
// of E the chunk needs is staged the same way (Ew), so the gather is 13 broadcast-friendly LDS.
// A particle whose stencil leaves the window (stale sort, sparse cells) takes the slow path: global
// atomics for the deposit, global loads for the gather; it is counted in slow_count.
// Window rows are Julia (1-based, unwrapped) indices base..base+31; wrap is applied at load/flush.
// ------------------------------------------------------------------------------------------
constexpr int WIN_ROWS = 32;                    // grid cells covered by a warp's window
constexpr int WIN_ALLOC = WIN_ROWS;             // out-of-window stencils are predicated off (slow path instead)
constexpr int WIN_LD = 32;                      // row stride: bank = lane for every row -> conflict-free whatever rows the lanes use
constexpr int WIN_LO = 6;                       // rows kept below the smallest centre of the first batch
constexpr int WIN_MAXOFF = WIN_ROWS - GAUSS_NW; // largest row a stencil may start at
constexpr int WIN_WARP_DOUBLES = WIN_ALLOC * WIN_LD + WIN_ROWS;
#ifndef PG_SORTED_NP
#define PG_SORTED_NP 2
#endif
constexpr int SORTED_NP = PG_SORTED_NP;         // particles evaluated together per lane

// Rare path: stencil outside the warp window.  Kept out of line so the hot loop stays small.
__device__ __noinline__ double slow_gather(const double *E, double c, int N)
{
    int ibase; double W[GAUSS_NW];
    gauss_weights(c, (double)N, ibase, W);
    double g = 0.0;
    for (int q = 0; q < GAUSS_NW; ++q) g = fma(__ldg(&E[(ibase + q - 1) & (N - 1)]), W[q], g);
    return g;
}
__device__ __noinline__ void slow_deposit(fx_t *rho, double c, int N, double fx_scale)
{
    int ibase; double W[GAUSS_NW];
    gauss_weights(c, (double)N, ibase, W);
    for (int q = 0; q < GAUSS_NW; ++q) atomicAdd(&rho[(ibase + q - 1) & (N - 1)], to_fx(W[q], fx_scale));
}

#ifndef PG_SORTED_MINBLOCKS
#define PG_SORTED_MINBLOCKS 2
#endif
template <bool FIRST, int NP>
__global__ void __launch_bounds__(PG_THREADS, PG_SORTED_MINBLOCKS) fp_pass_sorted(FPArgs a)
{
    extern __shared__ double smem[];
    __shared__ double scratch[32];
    const int fk = a.ctrl->final_k, k = FIRST ? 0 : sweep_index(a.k, a.ctrl);
    if (!FIRST && fk >= 0 && k > fk) return;
    const bool final = !FIRST && fk == k;
    const bool v0_is_V = FIRST || k == 1; // sweep 1 starts from v = V (`V.=v`): the work buffer is stale until pass 1 writes it
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    double *acc = smem + warp * WIN_WARP_DOUBLES; // [WIN_ALLOC][WIN_LD]
    double *Ew = acc + WIN_ALLOC * WIN_LD;        // [WIN_ROWS]
    const int N = a.N, Nmask = N - 1;
    const double dN = (double)N, dt = a.dt;
    const long long CH = 32LL * a.K; // a.K is a multiple of NP
    const long long nchunks = (a.P + CH - 1) / CH;
    double sv2 = 0.0, sv = 0.0;
    unsigned int nslow = 0;
    fx_t *const rho_out = final ? a.rho_next : a.rho; // the fused first deposit of the next step has its own grid
    for (long long ch = (long long)blockIdx.x * wpb + warp; ch < nchunks; ch += (long long)gridDim.x * wpb) {
        const long long j0 = ch * CH;
        // window base: smallest cell centre among the first batch of the chunk (same in all lanes)
        int base;
        {
            long long jf = j0 + lane < a.P ? j0 + lane : j0;
            int c0 = (int)rint(a.X[jf] * dN);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c0 = min(c0, __shfl_xor_sync(0xffffffffu, c0, o));
            base = c0 - 6 - WIN_LO;
        }
        if (!FIRST) Ew[lane] = a.E[(base + lane - 1) & Nmask];
#pragma unroll 4
        for (int r = 0; r < WIN_ROWS; ++r) acc[r * WIN_LD + lane] = 0.0;
        __syncwarp();
        // software pipeline: the loads of the next group are issued before the current group is evaluated
        double Xn[NP], Vn[NP], vn[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const long long jn = j0 + 32LL * q + lane;
            const bool ln = jn < a.P;
            Xn[q] = ln ? ld_stream(a.X + jn) : 0.0;
            Vn[q] = ln ? ld_stream(a.V + jn) : 0.0;
            vn[q] = v0_is_V ? Vn[q] : (ln ? ld_stream(a.v + jn) : 0.0);
        }
        for (int kb = 0; kb < a.K; kb += NP) {
            long long j[NP];
            bool live[NP];
            double Xj[NP], Vj[NP], vj[NP], xj[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                j[q] = j0 + 32LL * (kb + q) + lane;
                live[q] = j[q] < a.P;
                Xj[q] = Xn[q]; Vj[q] = Vn[q]; vj[q] = vn[q];
            }
            if (!live[0]) break;
            if (kb + NP < a.K) {
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const long long jn = j[q] + 32LL * NP;
                    const bool ln = jn < a.P;
                    Xn[q] = ln ? ld_stream(a.X + jn) : 0.0;
                    Vn[q] = ln ? ld_stream(a.V + jn) : 0.0;
                    vn[q] = v0_is_V ? Vn[q] : (ln ? ld_stream(a.v + jn) : 0.0);
                }
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) xj[q] = Xj[q] + (vj[q] + Vj[q]) / 2 * dt; // x.=X.+(v.+V)/2*dt
            if (!FIRST) {
                double d[NP], g[NP], mid[NP];
                const double *e[NP];
                bool ok[NP];
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    int ibase;
                    mid[q] = (xj[q] + Xj[q]) / 2;
                    gauss_centre(mid[q], dN, ibase, d[q]);
                    const int off = (ibase - base) & Nmask; // periodic: Julia indices 0 and N are the same cell
                    ok[q] = off <= WIN_MAXOFF;
                    e[q] = Ew + (ok[q] ? off : 0) + 6;
                    g[q] = 0.0;
                }
                gauss_stream<NP>(
                    d,
                    [&](auto m, const double(&wp)[NP], const double(&wm)[NP]) {
                        constexpr int M = decltype(m)::value;
#pragma unroll
                        for (int q = 0; q < NP; ++q) { g[q] = fma(e[q][M], wp[q], g[q]); g[q] = fma(e[q][-M], wm[q], g[q]); }
                    },
                    [&](const double(&w0)[NP]) {
#pragma unroll
                        for (int q = 0; q < NP; ++q) g[q] = fma(e[q][0], w0[q], g[q]);
                    });
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    if (!ok[q] && live[q]) g[q] = slow_gather(a.E, mid[q], N);
                    vj[q] = Vj[q] + g[q] * dt; // v[j]=V[j]+sum(...)*dt
                    if (live[q]) st_stream(a.v + j[q], vj[q]);
                }
                if (final) {
                    // end of step: x.=mod.(x,1), diagnostics sums -- and the first pass of the NEXT step fused in:
                    // X.=x; V.=v; x = X + (V+V)/2*dt; deposit at (x+X)/2, so the next step starts at its first solve
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const double xw = jl_mod1(xj[q]);
                        if (live[q]) {
                            st_stream(a.xout + j[q], xw);
                            sv2 = fma(vj[q], vj[q], sv2);
                            sv += vj[q];
                        }
                        Xj[q] = xw; Vj[q] = vj[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < NP; ++q) xj[q] = Xj[q] + (vj[q] + Vj[q]) / 2 * dt;
            }
            // deposit at (x+X)/2 into the lane-private window column
            double d[NP], mid[NP];
            double *col[NP];
            bool ok[NP], merged[NP], own[NP]; // merged[q]: summed into particle 0's rows in registers; own[q]: own RMW
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                int ibase;
                mid[q] = (xj[q] + Xj[q]) / 2;
                gauss_centre(mid[q], dN, ibase, d[q]);
                const int o = (ibase - base) & Nmask;
                ok[q] = live[q] && o <= WIN_MAXOFF;
                col[q] = acc + ((ok[q] ? o : 0) + 6) * WIN_LD + lane;
                merged[q] = q > 0 && ok[q] && ok[0] && col[q] == col[0];
                own[q] = ok[q] && !merged[q];
            }
            // One stencil evaluation for all NP particles, no divergent code paths: particles in the same rows as
            // particle 0 (the usual case in cell order) share one read-modify-write per cell; the others get a
            // predicated one of their own; out-of-window particles are predicated off and handled below.
            gauss_stream<NP>(
                d,
                [&](auto m, const double(&wp)[NP], const double(&wm)[NP]) {
                    constexpr int M = decltype(m)::value;
                    double sp = wp[0], sm = wm[0];
#pragma unroll
                    for (int q = 1; q < NP; ++q) { sp += merged[q] ? wp[q] : 0.0; sm += merged[q] ? wm[q] : 0.0; }
                    if (own[0]) { col[0][M * WIN_LD] += sp; col[0][-M * WIN_LD] += sm; }
#pragma unroll
                    for (int q = 1; q < NP; ++q)
                        if (own[q]) { col[q][M * WIN_LD] += wp[q]; col[q][-M * WIN_LD] += wm[q]; }
                },
                [&](const double(&w0)[NP]) {
                    double s0 = w0[0];
#pragma unroll
                    for (int q = 1; q < NP; ++q) s0 += merged[q] ? w0[q] : 0.0;
                    if (own[0]) col[0][0] += s0;
#pragma unroll
                    for (int q = 1; q < NP; ++q)
                        if (own[q]) col[q][0] += w0[q];
                });
#pragma unroll
            for (int q = 0; q < NP; ++q)
                if (live[q] && !ok[q]) { ++nslow; slow_deposit(rho_out, mid[q], N, a.fx_scale); }
        }
        __syncwarp();
        {
            double s = 0.0;
            const double *row = acc + lane * WIN_LD;
#pragma unroll 8
            for (int c = 0; c < 32; ++c) s += row[(c + lane) & 31]; // skewed: lane r starts at column r (distinct banks)
            if (s != 0.0) atomicAdd(&rho_out[(base + lane - 1) & Nmask], to_fx(s, a.fx_scale)); // one integer RED per window cell
        }
        __syncwarp();
    }
    if (final) {
        sv2 = block_sum(sv2, scratch);
        sv = block_sum(sv, scratch);
        if (threadIdx.x == 0) { a.partials[2 * blockIdx.x] = sv2; a.partials[2 * blockIdx.x + 1] = sv; }
    }
    if (nslow && a.slow_count) atomicAdd(a.slow_count, (unsigned long long)nslow);
}

// ------------------------------------------------------------------------------------------
// Leapfrog passes (NGP and explicit Gaussian).  One pass = [second half drift + kick of step t]
// followed by [first half drift + deposit of step t+1]; either half can be switched off so the
// state handed back to the caller is always an end-of-step state.
//   u():  x .= mod.(x .+ v/2*dt, 1)        NGPFourier.jl:2
//   kick: v += E[f.(x)]*dt                 NGPFourier.jl:6 ;  Gaussian.jl:10
//   deposit: n[f(j)] += w                  NGPFourier.jl:5 ;  Gaussian.jl:7
// NGP deposits are exact integer counts (shared-memory u32, global u64); rho = count*w is formed by
// the solve kernel, so NGP charge is bit-reproducible and independent of particle order.  The explicit
// Gaussian deposits in fixed point like the fixed-point scheme above.
// ------------------------------------------------------------------------------------------
struct LFArgs {
    double *x, *v;
    const double *E;
    fx_t *rho;                   // integer deposit grid: NGP counts (frac = 0) or Gaussian fixed point
    double *partials;            // [2*gridDim.x]
    long long P;
    double dt, fx_scale;
    int N, do_kick, do_deposit;  // do_deposit = 2: deposit at the half-drifted position but store the full-step x (last step of a call)
    int pow2;                    // N is a power of two: the cell index is a mask instead of Julia's mod1 division
    int do_predrift;             // the stored x is a full-step position whose half drift u() was only deposited (do_deposit = 2 of the
                                 // previous call): redo it first -- the same operands, the same bits
};

// One particle of a leapfrog pass (shared by the vectorised and the scalar loops).  EDGE: the first / last pass of a picgolf_step call
// (do_predrift, do_deposit = 2) -- a separate instantiation, so that the passes in between carry none of it (at run time the extra
// mod and select cost the TMA-staged NGP pass 5 %).
template <int SHAPE, bool EDGE>
__device__ __forceinline__ void lf_particle(const LFArgs &a, const double *Es, fx_t *rs, unsigned int *cs, double &xj, double &vj,
                                            double &sv2, double &sv)
{
    const int N = a.N, Nmask = N - 1;
    const double dN = (double)N, dt = a.dt;
    if (EDGE && a.do_predrift) xj = jl_mod1(xj + vj / 2 * dt);
    if (a.do_kick) {
        xj = jl_mod1(xj + vj / 2 * dt);
        double e;
        if (SHAPE == 0) e = Es[a.pow2 ? ngp_cell0_pow2(xj, N) : ngp_cell0(xj, N)];
        else {
            int ibase; double W[GAUSS_NW];
            gauss_weights(xj, dN, ibase, W);
            e = gauss_gather(Es, ibase, W, Nmask);
        }
        vj = vj + e * dt;
        sv2 = fma(vj, vj, sv2);
        sv += vj;
    }
    if (a.do_deposit) {
        const double xd = jl_mod1(xj + vj / 2 * dt);
        if (SHAPE == 0) atomicAdd(&cs[a.pow2 ? ngp_cell0_pow2(xd, N) : ngp_cell0(xd, N)], 1u);
        else {
            int ibase; double W[GAUSS_NW];
            gauss_weights(xd, dN, ibase, W);
            gauss_deposit_atomic(rs, ibase, W, a.fx_scale, Nmask);
        }
        if (!EDGE || a.do_deposit == 1) xj = xd;
    }
}

// Shared memory: Es[N] doubles, then the block-private deposit grid (NGP: N u32 counts; Gaussian: N
// int64 fixed point), then 32 doubles of reduction scratch.  NGP streams the particle arrays with 128-bit
// loads/stores, two pairs (4 particles) in flight per thread.
__host__ __device__ inline size_t lf_smem_bytes(int shape, int N) { return (size_t)N * 8 + (size_t)N * (shape == 0 ? 4 : 8) + 256; }

template <int SHAPE, bool EDGE> // 0 = NGP, 1 = Gaussian
__global__ void __launch_bounds__(PG_THREADS) lf_pass(LFArgs a)
{
    extern __shared__ double smem[];
    double *Es = smem;
    fx_t *rs = reinterpret_cast<fx_t *>(smem + a.N);
    unsigned int *cs = reinterpret_cast<unsigned int *>(smem + a.N);
    double *scratch = reinterpret_cast<double *>(reinterpret_cast<char *>(smem) + lf_smem_bytes(SHAPE, a.N) - 256);
    const int N = a.N;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        if (a.do_kick) Es[n] = a.E[n];
        if (SHAPE == 0) cs[n] = 0u; else rs[n] = 0ULL;
    }
    __syncthreads();
    double sv2 = 0.0, sv = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (SHAPE == 0) {
        const long long npair = a.P >> 1;
        double2 *x2 = reinterpret_cast<double2 *>(a.x), *v2 = reinterpret_cast<double2 *>(a.v);
        long long p = tid;
        for (; p + stride < npair; p += 2 * stride) { // two 128-bit loads per array in flight
            double2 xa = __ldcs(x2 + p), va = __ldcs(v2 + p), xb = __ldcs(x2 + p + stride), vb = __ldcs(v2 + p + stride);
            lf_particle<SHAPE, EDGE>(a, Es, rs, cs, xa.x, va.x, sv2, sv);
            lf_particle<SHAPE, EDGE>(a, Es, rs, cs, xa.y, va.y, sv2, sv);
            lf_particle<SHAPE, EDGE>(a, Es, rs, cs, xb.x, vb.x, sv2, sv);
            lf_particle<SHAPE, EDGE>(a, Es, rs, cs, xb.y, vb.y, sv2, sv);
            __stcs(x2 + p, xa); __stcs(v2 + p, va); __stcs(x2 + p + stride, xb); __stcs(v2 + p + stride, vb);
        }
        for (; p < npair; p += stride) {
            double2 xa = __ldcs(x2 + p), va = __ldcs(v2 + p);
            lf_particle<SHAPE, EDGE>(a, Es, rs, cs, xa.x, va.x, sv2, sv);
            lf_particle<SHAPE, EDGE>(a, Es, rs, cs, xa.y, va.y, sv2, sv);
            __stcs(x2 + p, xa); __stcs(v2 + p, va);
        }
        if ((a.P & 1) && tid == 0) { // odd tail
            double xj = a.x[a.P - 1], vj = a.v[a.P - 1];
            lf_particle<SHAPE, EDGE>(a, Es, rs, cs, xj, vj, sv2, sv);
            a.x[a.P - 1] = xj; a.v[a.P - 1] = vj;
        }
    } else {
        for (long long j = tid; j < a.P; j += stride) {
            double xj = ld_stream(a.x + j), vj = ld_stream(a.v + j);
            lf_particle<SHAPE, EDGE>(a, Es, rs, cs, xj, vj, sv2, sv);
            st_stream(a.x + j, xj);
            st_stream(a.v + j, vj);
        }
    }
    __syncthreads();
    if (a.do_deposit) {
        if (SHAPE == 0) {
            for (int n = threadIdx.x; n < N; n += blockDim.x) {
                unsigned int c = cs[n];
                if (c) atomicAdd(&a.rho[n], (fx_t)c);
            }
        } else flush_grid(rs, a.rho, N);
    }
    if (a.do_kick) {
        sv2 = block_sum(sv2, scratch);
        sv = block_sum(sv, scratch);
        if (threadIdx.x == 0) { a.partials[2 * blockIdx.x] = sv2; a.partials[2 * blockIdx.x + 1] = sv; }
    }
}

// ------------------------------------------------------------------------------------------
// NGP leapfrog pass with TMA-staged particle tiles (the HBM-bound path: 32 B per particle-step).
// One block per SM streams tiles of LF_TILE particles through a LF_STAGES-deep shared-memory ring:
//   thread 0 issues `cp.async.bulk` loads of the x and v tiles (completion on an mbarrier), all threads
//   update the tile in place in shared memory (lf_particle, identical arithmetic to lf_pass<0>), then
//   thread 0 writes the tile back with bulk stores and refills the stage that was stored one iteration ago.
// The copy engine moves all particle bytes; the SM only touches shared memory.  Full tiles only; the
// remainder (< LF_TILE particles) is handled by the scalar tail below.
// ------------------------------------------------------------------------------------------
constexpr int LF_TILE = 2048;  // particles per tile: 16 KB of x + 16 KB of v
constexpr int LF_STAGES = 3;
__host__ __device__ inline size_t lf_tma_smem_bytes(int N)
{
    return (size_t)LF_STAGES * LF_TILE * 16 + (size_t)N * 8 + (size_t)N * 4 + 256 + 64; // ring, Es, counts, scratch, barriers
}

constexpr int LF_TMA_THREADS = 512;
template <bool EDGE>
__global__ void __launch_bounds__(LF_TMA_THREADS, 1) lf_pass_ngp_tma(LFArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);                    // [STAGES][2][TILE]
    double *Es = ring + (size_t)LF_STAGES * 2 * LF_TILE;
    unsigned int *cs = reinterpret_cast<unsigned int *>(Es + a.N);
    double *scratch = reinterpret_cast<double *>(cs + a.N);
    uint64_t *full = reinterpret_cast<uint64_t *>(scratch + 32);            // [STAGES]
    const int N = a.N;
    const long long ntiles = a.P / LF_TILE;
    const long long mine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0; // tiles of this block
    constexpr uint32_t TILE_BYTES = LF_TILE * 8;
    if (threadIdx.x == 0) {
        for (int s = 0; s < LF_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        if (a.do_kick) Es[n] = a.E[n];
        cs[n] = 0u;
    }
    __syncthreads();
    auto issue_load = [&](long long it) { // tile number `it` of this block into stage it % STAGES
        const int s = (int)(it % LF_STAGES);
        const long long j0 = ((long long)blockIdx.x + it * gridDim.x) * LF_TILE;
        double *xs = ring + (size_t)s * 2 * LF_TILE, *vs = xs + LF_TILE;
        mbar_arrive_expect_tx(&full[s], 2 * TILE_BYTES);
        bulk_load(xs, a.x + j0, TILE_BYTES, &full[s]);
        bulk_load(vs, a.v + j0, TILE_BYTES, &full[s]);
    };
    if (threadIdx.x == 0)
        for (long long it = 0; it < LF_STAGES - 1 && it < mine; ++it) issue_load(it);
    double sv2 = 0.0, sv = 0.0;
    for (long long it = 0; it < mine; ++it) {
        const int s = (int)(it % LF_STAGES);
        double *xs = ring + (size_t)s * 2 * LF_TILE, *vs = xs + LF_TILE;
        mbar_wait(&full[s], (uint32_t)((it / LF_STAGES) & 1));
#pragma unroll
        for (int i = 0; i < LF_TILE / LF_TMA_THREADS; i += 2) { // pairs via 128-bit shared accesses
            const int p = (i / 2) * LF_TMA_THREADS + threadIdx.x;
            double2 xa = reinterpret_cast<double2 *>(xs)[p], va = reinterpret_cast<double2 *>(vs)[p];
            lf_particle<0, EDGE>(a, Es, nullptr, cs, xa.x, va.x, sv2, sv);
            lf_particle<0, EDGE>(a, Es, nullptr, cs, xa.y, va.y, sv2, sv);
            reinterpret_cast<double2 *>(xs)[p] = xa; reinterpret_cast<double2 *>(vs)[p] = va;
        }
        fence_proxy_async(); // the updated tile must be visible to the copy engine
        __syncthreads();
        if (threadIdx.x == 0) {
            const long long j0 = ((long long)blockIdx.x + it * gridDim.x) * LF_TILE;
            bulk_store(a.x + j0, xs, TILE_BYTES);
            bulk_store(a.v + j0, vs, TILE_BYTES);
            bulk_commit();
            const long long nxt = it + LF_STAGES - 1; // goes into the stage stored one iteration ago
            if (nxt < mine) {
                bulk_wait_read<1>(); // ... whose store must have finished reading shared memory
                issue_load(nxt);
            }
        }
    }
    if (threadIdx.x == 0) bulk_wait_all<0>();
    // scalar tail: the last P % LF_TILE particles, spread over the blocks
    {
        const long long t0 = ntiles * LF_TILE;
        for (long long j = t0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.P; j += (long long)gridDim.x * blockDim.x) {
            double xj = a.x[j], vj = a.v[j];
            lf_particle<0, EDGE>(a, Es, nullptr, cs, xj, vj, sv2, sv);
            a.x[j] = xj; a.v[j] = vj;
        }
    }
    __syncthreads();
    if (a.do_deposit)
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            unsigned int c = cs[n];
            if (c) atomicAdd(&a.rho[n], (fx_t)c);
        }
    if (a.do_kick) {
        sv2 = block_sum(sv2, scratch);
        sv = block_sum(sv, scratch);
        if (threadIdx.x == 0) { a.partials[2 * blockIdx.x] = sv2; a.partials[2 * blockIdx.x + 1] = sv; }
    }
}

// ------------------------------------------------------------------------------------------
// 1D2V magnetised Boris pass, src/NGP1D2V.jl:41-45,55: gather E at x with the erf shape,
// boris() about z (:5-10), x += vx*dt, x = mod(x,1); fused with the deposit rho(x) of the next step (:40).
// 48 B per particle-step (x, vx, vy read + written).  Any particle order; fixed-point deposit grid.
// ------------------------------------------------------------------------------------------
// Two-species variant, src/NGP1D2V2S.jl:5-11,24-25,31-52: the particle arrays hold species 1 (global indices < Psp:
// q = -1, q/m = -1) followed by species 2 (q = +1, q/m = 1/M); boris(vx,vy,E,B,dt,q_m) scales E and B by
// dt2q_m = dt/2*q_m; the deposit is r[k[1]] += q*k[2]*w; the kinetic sums weigh species 2 with its mass M.
struct B1D2VArgs {
    double *x, *vx, *vy;
    const double *E;
    fx_t *rho;
    double *partials; // [3*gridDim.x] per-block (sum m(vx^2+vy^2), sum m vx, sum m vy), m = 1 or M
    long long P;
    double dt, B0, fx_scale;
    double M;         // mass ratio of species 2 (two-species scheme)
    long long first;  // global index of local particle 0
    long long Psp;    // particles per species; < 0: single species (NGP1D2V.jl: q = 1, q/m = 1)
    int N, do_push, do_deposit;
};

__global__ void __launch_bounds__(PG_THREADS) b1d2v_pass(B1D2VArgs a)
{
    extern __shared__ double smem[];
    const int N = a.N, Nmask = N - 1;
    double *Es = smem, *scratch = smem + 2 * N;
    fx_t *rs = reinterpret_cast<fx_t *>(smem + N);
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        if (a.do_push) Es[n] = a.E[n];
        rs[n] = 0ULL;
    }
    __syncthreads();
    const double dN = (double)N, dt = a.dt;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.P; j += stride) {
        double xj = ld_stream(a.x + j), vx = ld_stream(a.vx + j), vy = ld_stream(a.vy + j);
        const bool sp2 = a.Psp >= 0 && a.first + j >= a.Psp;           // second species of NGP1D2V2S.jl
        const double q = (a.Psp >= 0 && !sp2) ? -1.0 : 1.0;
        const double q_m = a.Psp < 0 ? 1.0 : (sp2 ? 1 / a.M : -1.0);
        const double mass = sp2 ? a.M : 1.0;
        int ibase; double W[GAUSS_NW];
        if (a.do_push) {
            gauss_weights(xj, dN, ibase, W);
            const double Ej = gauss_gather(Es, ibase, W, Nmask);      // sum(k->E[k[1]]*k[2], d(x[j]))  :42
            // dt2q_m = dt/2*q_m (NGP1D2V2S.jl:6); with q_m = 1 the products below are bit-identical to the
            // E*dt/2 and B*dt/2 of NGP1D2V.jl:6-7 (halving is exact)
            const double hq = dt / 2 * q_m;
            const double h = Ej * hq, t3 = a.B0 * hq;
            const double den = 1 + (0.0 + 0.0 + t3 * t3);
            const double m1 = vx + h, m2 = vy;                         // v- = [vx + E*dt2q_m, vy, 0]
            const double p1 = m1 + m2 * t3, p2 = m2 - m1 * t3;         // v- + cross(v-, t), t = [0,0,B*dt2q_m]
            const double r1 = m1 + 2 * (p2 * t3) / den;                // v+ = v- + 2*cross(.., t)/(1+dot(t,t))
            const double r2 = m2 - 2 * (p1 * t3) / den;
            vx = r1 + h; vy = r2;
            xj = jl_mod1(xj + vx * dt);                                // x[j] += vx[j]*dt ; x.=mod.(x,1)
            s0 += mass * (vy * vy + vx * vx); s1 += mass * vx; s2 += mass * vy;
        }
        if (a.do_deposit) {
            gauss_weights((xj + xj) / 2, dN, ibase, W);                // rho(x)
#pragma unroll
            for (int k = 0; k < GAUSS_NW; ++k) W[k] *= q;              // r[k[1]] += q*k[2]*w   NGP1D2V2S.jl:24
            gauss_deposit_atomic(rs, ibase, W, a.fx_scale, Nmask);
        }
        st_stream(a.x + j, xj); st_stream(a.vx + j, vx); st_stream(a.vy + j, vy);
    }
    __syncthreads();
    if (a.do_deposit) flush_grid(rs, a.rho, N);
    if (a.do_push) {
        s0 = block_sum(s0, scratch);
        s1 = block_sum(s1, scratch);
        s2 = block_sum(s2, scratch);
        if (threadIdx.x == 0) { a.partials[3 * blockIdx.x] = s0; a.partials[3 * blockIdx.x + 1] = s1; a.partials[3 * blockIdx.x + 2] = s2; }
    }
}

// ------------------------------------------------------------------------------------------
// End of step: reduce the per-block partial sums in a fixed order, append the raw diagnostics row
// (column-major, ld = T): [sum E^2, sum v^2 (2D: vx^2+vy^2), sum v (vx), sweeps (2D: sum vy)].
// ------------------------------------------------------------------------------------------
struct StepEndArgs {
    const double *partials; // [npart * nblocks]
    const double *epartials; // optional per-block partials of sum(E^2) (2D); NULL -> ctrl->sumE2
    double *raw;            // [4*T]
    Ctrl *ctrl;
    int nblocks, npart, neblocks, T, record, is2d;
    int det; // deterministic polynomial passes: partials holds four 64-bit integers per block (pg_kernels_poly.cuh)
};

__global__ void __launch_bounds__(256) step_end_kernel(StepEndArgs a)
{
    __shared__ double scratch[32];
    double s[3] = {0.0, 0.0, 0.0};
    if (a.det) { // exact integer sums: the same bits whatever particles each block summed
        __shared__ long long isum[4];
        if (threadIdx.x < 4) isum[threadIdx.x] = 0;
        __syncthreads();
        const long long *pp = reinterpret_cast<const long long *>(a.partials);
        long long q[4] = {0, 0, 0, 0};
        for (int b = threadIdx.x; b < a.nblocks; b += blockDim.x)
            for (int c = 0; c < 4; ++c) q[c] += pp[4 * b + c];
        for (int c = 0; c < 4; ++c) atomicAdd(reinterpret_cast<unsigned long long *>(&isum[c]), (unsigned long long)q[c]);
        __syncthreads();
        if (threadIdx.x == 0) {
            const double sc = 1.0 / (double)(1LL << 40); // CP_DETV_FRAC
            s[0] = ((double)isum[1] * 4294967296.0 + (double)isum[0]) * sc;
            s[1] = ((double)isum[3] * 4294967296.0 + (double)isum[2]) * sc;
        }
    } else
    for (int b = threadIdx.x; b < a.nblocks; b += blockDim.x)
        for (int c = 0; c < a.npart; ++c) s[c] += a.partials[a.npart * b + c];
    double e = 0.0;
    if (a.epartials) for (int b = threadIdx.x; b < a.neblocks; b += blockDim.x) e += a.epartials[b];
    for (int c = 0; c < 3; ++c) s[c] = block_sum(s[c], scratch);
    e = block_sum(e, scratch);
    if (threadIdx.x == 0) {
        Ctrl *c = a.ctrl;
        if (a.epartials) c->sumE2 = e;
        if (a.record && c->rows < a.T) {
            int r = c->rows;
            a.raw[r] = c->sumE2;
            a.raw[a.T + r] = s[0];
            a.raw[2 * a.T + r] = s[1];
            a.raw[3 * a.T + r] = a.is2d ? s[2] : (double)c->sweeps;
            c->rows = r + 1;
        }
        c->step += 1;
        c->final_k = -1;
        c->fs_pred = c->last_sweeps > 0 ? min(c->last_sweeps, c->sweeps) : c->sweeps;
        c->last_sweeps = c->sweeps;
        c->sweeps = 0;
    }
}

// ------------------------------------------------------------------------------------------
// Initialisation
// ------------------------------------------------------------------------------------------
// x=(bitreverse.(0:P-1).+2.0^63)/2.0^64 (signed reinterpretation); v = (j > P/2) ? 1 : -1 (1-based j)
// src/GaussianFixedPointQuiet.jl:2-3.  `first` is the global index of local particle 0.
__global__ void quiet_start_kernel(double *x, double *v, long long count, long long first, long long P)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < count; n += stride) {
        long long s = (long long)__brevll((unsigned long long)(first + n));
        x[n] = ((double)s + 9223372036854775808.0) / 18446744073709551616.0;
        v[n] = ((double)(first + n + 1) > (double)P / 2) ? 1.0 : -1.0;
    }
}

// Seeded synthetic two-stream start (the NGPFourier.jl:2 pattern with a counter-based generator).  vth > 0 warms the
// beams: v = +-1 + vth * N(0,1) (Box-Muller on two more draws) -- the velocity spread of the saturated two-stream state
// that the cold start only reaches after ~10 plasma periods (bench.py's warm-regime line).
__global__ void synthetic_1d_kernel(double *x, double *v, long long count, long long first, long long P, uint64_t seed, double vth)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < count; n += stride) {
        uint64_t g = (uint64_t)(first + n);
        x[n] = u01(splitmix64(seed ^ (g * 0xD1342543DE82EF95ULL)));
        double vn = ((double)(first + n + 1) > (double)P / 2) ? 1.0 : -1.0;
        if (vth > 0.0) {
            const double u1 = u01(splitmix64(seed ^ ((g * 4 + 1) * 0x9E3779B97F4A7C15ULL))), u2 = u01(splitmix64(seed ^ ((g * 4 + 2) * 0xC2B2AE3D27D4EB4FULL)));
            vn += vth * sqrt(-2.0 * log(1.0 - u1)) * cospi(2.0 * u2);
        }
        v[n] = vn;
    }
}

// ------------------------------------------------------------------------------------------
// Stage-level kernels behind picgolf_stage_* (parity tests exercise the same device functions)
// ------------------------------------------------------------------------------------------
__global__ void stage_ngp_index_kernel(const double *x, long long count, int N, int *idx1)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < count) idx1[j] = ngp_cell0(x[j], N) + 1;
}

__global__ void stage_mod1_kernel(const double *x, long long count, double *out)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < count) out[j] = jl_mod1(x[j]);
}

__global__ void stage_gauss_stencil_kernel(const double *c, long long count, int N, int hw, int *idx1, double *wt)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    int ibase; double W[GAUSS_NW];
    gauss_weights(c[j], (double)N, ibase, W);
    const int nw = 2 * hw + 1, pad = hw - 6;
    for (int k = 0; k < nw; ++k) {
        int i = ibase - pad + k; // Julia index of offset -hw+k
        idx1[j * nw + k] = wrap_cell0(i, N) + 1;
        int kk = k - pad;
        double val = 0.0;
#pragma unroll
        for (int q = 0; q < GAUSS_NW; ++q) if (q == kk) val = W[q];
        wt[j * nw + k] = val;
    }
}

__global__ void __launch_bounds__(PG_THREADS) stage_gauss_deposit_kernel(const double *x, const double *y, long long count,
                                                                        int N, double scale, fx_t *rho)
{
    extern __shared__ double smem[];
    fx_t *rs = reinterpret_cast<fx_t *>(smem);
    for (int n = threadIdx.x; n < N; n += blockDim.x) rs[n] = 0ULL;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
        int ibase; double W[GAUSS_NW];
        gauss_weights((x[j] + y[j]) / 2, (double)N, ibase, W);
        gauss_deposit_atomic(rs, ibase, W, scale, N - 1);
    }
    __syncthreads();
    flush_grid(rs, rho, N);
}

__global__ void stage_gauss_gather_kernel(const double *E, int N, const double *c, long long count, double *out)
{
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    int ibase; double W[GAUSS_NW];
    gauss_weights(c[j], (double)N, ibase, W);
    out[j] = gauss_gather(E, ibase, W, N - 1);
}

} // namespace pg
