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

namespace pg {

// ------------------------------------------------------------------------------------------
// deposit primitives on a block-private shared-memory grid
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void gauss_deposit_atomic(double *rs, int ibase, const double (&W)[GAUSS_NW], double scale,
                                                     int Nmask)
{
    // r[k[1]] += k[2]*w      src/GaussianFixedPoint.jl:6
#pragma unroll
    for (int k = 0; k < GAUSS_NW; ++k) atomicAdd(&rs[(ibase + k - 1) & Nmask], W[k] * scale);
}

__device__ __forceinline__ void flush_grid(const double *rs, double *rho, int N)
{
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double r = rs[n];
        if (r != 0.0) atomicAdd(&rho[n], r);
    }
}

// ------------------------------------------------------------------------------------------
// Gaussian fixed point.  Pass k of a step (k = 0..max_sweeps):
//   k = 0 (FIRST): x_0 = X + ((V+V)/2)*dt; deposit at (x_0+X)/2.
//   k >= 1: gather E_k at (x_{k-1}+X)/2 -> v_k = V + g*dt.  If the solve that produced E_k declared
//           the step finished (ctrl->final_k == k): x = mod(x_{k-1},1), accumulate sum v^2, sum v.
//           Otherwise x_k = X + ((v_k+V)/2)*dt and deposit at (x_k+X)/2 for solve k+1.
//   Passes with k > final_k are predicated no-ops (the host never reads the flag mid-step).
// Note x lags v by one sweep on exit, exactly as in the reference (GaussianFixedPoint.jl:8-9).
// ------------------------------------------------------------------------------------------
struct FPArgs {
    const double *X, *V; // step-start state (X.=x; V.=v)
    double *v;           // working velocity v_k (in place)
    double *xout;        // wrapped end-of-step position
    const double *E;     // field of solve k
    double *rho;         // global deposit grid (zeroed by the solve that consumed it)
    double *partials;    // [2*gridDim.x] per-block (sum v^2, sum v)
    Ctrl *ctrl;
    long long P;
    double dt, w;
    int N, k;
};

template <bool FIRST>
__global__ void __launch_bounds__(PG_THREADS) fp_pass_atomic(FPArgs a)
{
    extern __shared__ double smem[];
    double *Es = smem, *rs = smem + a.N, *scratch = smem + 2 * a.N;
    const int fk = a.ctrl->final_k;
    if (!FIRST && fk >= 0 && a.k > fk) return;
    const bool final = !FIRST && fk == a.k;
    const int N = a.N, Nmask = N - 1;
    const double dN = (double)N, dt = a.dt;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        if (!FIRST) Es[n] = a.E[n];
        rs[n] = 0.0;
    }
    __syncthreads();
    double sv2 = 0.0, sv = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.P; j += stride) {
        const double Xj = ld_stream(a.X + j), Vj = ld_stream(a.V + j);
        double vj = FIRST ? Vj : ld_stream(a.v + j);
        double xj = Xj + (vj + Vj) / 2 * dt; // x.=X.+(v.+V)/2*dt
        int ibase;
        double W[GAUSS_NW];
        if (!FIRST) {
            gauss_weights((xj + Xj) / 2, dN, ibase, W);
            double g = gauss_gather(Es, ibase, W, Nmask);
            vj = Vj + g * dt; // v[j]=V[j]+sum(...)*dt
            st_stream(a.v + j, vj);
            if (final) {
                st_stream(a.xout + j, jl_mod1(xj)); // x.=mod.(x,1)
                sv2 = fma(vj, vj, sv2);
                sv += vj;
                continue;
            }
            xj = Xj + (vj + Vj) / 2 * dt;
        }
        gauss_weights((xj + Xj) / 2, dN, ibase, W);
        gauss_deposit_atomic(rs, ibase, W, a.w, Nmask);
    }
    __syncthreads();
    if (final) {
        sv2 = block_sum(sv2, scratch);
        sv = block_sum(sv, scratch);
        if (threadIdx.x == 0) { a.partials[2 * blockIdx.x] = sv2; a.partials[2 * blockIdx.x + 1] = sv; }
    } else {
        flush_grid(rs, a.rho, N);
    }
}

// ------------------------------------------------------------------------------------------
// Leapfrog passes (NGP and explicit Gaussian).  One pass = [second half drift + kick of step t]
// followed by [first half drift + deposit of step t+1]; either half can be switched off so the
// state handed back to the caller is always an end-of-step state.
//   u():  x .= mod.(x .+ v/2*dt, 1)        NGPFourier.jl:2
//   kick: v += E[f.(x)]*dt                 NGPFourier.jl:6 ;  Gaussian.jl:10
//   deposit: n[f(j)] += w                  NGPFourier.jl:5 ;  Gaussian.jl:7
// NGP deposits are exact integer counts (shared-memory u32, global u64); rho = count*w is formed by
// the solve kernel, so NGP charge is bit-reproducible and independent of particle order.
// ------------------------------------------------------------------------------------------
struct LFArgs {
    double *x, *v;
    const double *E;
    double *rho;                 // Gaussian
    unsigned long long *counts;  // NGP
    double *partials;            // [2*gridDim.x]
    long long P;
    double dt, w;
    int N, do_kick, do_deposit;
};

template <int SHAPE> // 0 = NGP, 1 = Gaussian
__global__ void __launch_bounds__(PG_THREADS) lf_pass(LFArgs a)
{
    extern __shared__ double smem[];
    double *Es = smem, *rs = smem + a.N, *scratch = smem + 2 * a.N;
    unsigned int *cs = reinterpret_cast<unsigned int *>(rs);
    const int N = a.N, Nmask = N - 1;
    const double dN = (double)N, dt = a.dt;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        if (a.do_kick) Es[n] = a.E[n];
        if (SHAPE == 0) cs[n] = 0u; else rs[n] = 0.0;
    }
    __syncthreads();
    double sv2 = 0.0, sv = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.P; j += stride) {
        double xj = ld_stream(a.x + j), vj = ld_stream(a.v + j);
        if (a.do_kick) {
            xj = jl_mod1(xj + vj / 2 * dt);
            double e;
            if (SHAPE == 0) e = Es[ngp_cell0(xj, N)];
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
            xj = jl_mod1(xj + vj / 2 * dt);
            if (SHAPE == 0) atomicAdd(&cs[ngp_cell0(xj, N)], 1u);
            else {
                int ibase; double W[GAUSS_NW];
                gauss_weights(xj, dN, ibase, W);
                gauss_deposit_atomic(rs, ibase, W, a.w, Nmask);
            }
        }
        st_stream(a.x + j, xj);
        st_stream(a.v + j, vj);
    }
    __syncthreads();
    if (a.do_deposit) {
        if (SHAPE == 0) {
            for (int n = threadIdx.x; n < N; n += blockDim.x) {
                unsigned int c = cs[n];
                if (c) atomicAdd(&a.counts[n], (unsigned long long)c);
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
// End of step: reduce the per-block partial sums in a fixed order, append the raw diagnostics row
// (column-major, ld = T): [sum E^2, sum v^2 (2D: vx^2+vy^2), sum v (vx), sweeps (2D: sum vy)].
// ------------------------------------------------------------------------------------------
struct StepEndArgs {
    const double *partials; // [npart * nblocks]
    const double *epartials; // optional per-block partials of sum(E^2) (2D); NULL -> ctrl->sumE2
    double *raw;            // [4*T]
    Ctrl *ctrl;
    int nblocks, npart, neblocks, T, record, is2d;
};

__global__ void __launch_bounds__(256) step_end_kernel(StepEndArgs a)
{
    __shared__ double scratch[32];
    double s[3] = {0.0, 0.0, 0.0};
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

// Seeded synthetic two-stream start (the NGPFourier.jl:2 pattern with a counter-based generator).
__global__ void synthetic_1d_kernel(double *x, double *v, long long count, long long first, long long P, uint64_t seed)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < count; n += stride) {
        uint64_t g = (uint64_t)(first + n);
        x[n] = u01(splitmix64(seed ^ (g * 0xD1342543DE82EF95ULL)));
        v[n] = ((double)(first + n + 1) > (double)P / 2) ? 1.0 : -1.0;
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
                                                                        int N, double scale, double *rho)
{
    extern __shared__ double smem[];
    double *rs = smem;
    for (int n = threadIdx.x; n < N; n += blockDim.x) rs[n] = 0.0;
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
