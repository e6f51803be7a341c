// pg_kernels_simpson.cuh -- Simpson-1/3 fixed-point variant, src/GaussianFixedPointQuietSimpson13.jl:8-18
// (same deposit/solve/gather kernels, new sweep schedule).  Per step:
//     E1 = solve(rho(X,X))                                                     :9
//     sweep k: x_k = X + (v_{k-1}+V)/2*dt                                      :11
//              v   = V + g(E1, X)*dt/6                                         :12
//              E2  = solve(rho(X, x_k));  v += g(E2,(X+x_k)/2)*4dt/6           :13-14
//              E3  = solve(rho(x_k,x_k)); v += g(E3, x_k)*dt/6                 :15-16
//     until isapprox(F, E) on the whole 3 x N field matrix (Frobenius norm)     :10
// Passes (any particle order; block-private shared-memory grids, fixed-point deposits):
//     sp_pass0          deposit at X                    -> rho1            -> solve E1
//     sp_pass1          g1 = g(E1, X) (constant over the sweeps, parked in the idle xout buffer),
//                       x_1 = X + V*dt, deposit at (X+x_1)/2 and at x_1   -> rho2, rho3 -> solve E2, E3
//     sp_passk (k>=1)   gather E2, E3 -> v_k; final: x = mod(x_k,1), sums; else x_{k+1} and the two deposits
// E and rho hold three rows of N: E1|E2|E3, rho1|rho2|rho3.
#pragma once
#include "pg_common.cuh"
#include "pg_fft.cuh"
#include "pg_gauss.cuh"
#include "pg_kernels_1d.cuh"

namespace pg {

// Particle shape of the Simpson-1/3 scripts: SHAPE 0 = erf-integrated Gaussian (13 cells,
// GaussianFixedPointQuietSimpson13.jl:5-6), SHAPE 1 = "area"/CIC (2 cells, AreaFixedPointQuietSimpson13.jl:5:
//     d(y)=(i=Int(mod1(ceil(y*N),N)); o=ceil(y*N)-y*N; ((i,1-o),(mod1(i-1,N),o)))  ).
// Entry k of a stencil is the Julia (1-based, unwrapped) cell ibase+k; cell0 = (ibase+k-1) & (N-1).
template <int SHAPE> struct SpShape;
template <> struct SpShape<0> {
    static constexpr int NW = GAUSS_NW;
    __device__ static __forceinline__ void weights(double c, double dN, int &ibase, double (&W)[NW]) { gauss_weights(c, dN, ibase, W); }
};
template <> struct SpShape<1> {
    static constexpr int NW = 2;
    __device__ static __forceinline__ void weights(double y, double dN, int &ibase, double (&W)[NW])
    {
        double yN = y * dN, ce = ceil(yN);
        double o = ce - yN;
        ibase = (int)ce - 1; // entries: (i-1, o), (i, 1-o)
        W[0] = o; W[1] = 1 - o;
    }
};
template <int NW>
__device__ __forceinline__ double sp_gather(const double *Es, int ibase, const double (&W)[NW], int Nmask)
{
    double g = 0.0;
#pragma unroll
    for (int k = 0; k < NW; ++k) g = fma(Es[(ibase + k - 1) & Nmask], W[k], g);
    return g;
}
template <int NW>
__device__ __forceinline__ void sp_deposit(fx_t *rs, int ibase, const double (&W)[NW], double fx_scale, int Nmask)
{
#pragma unroll
    for (int k = 0; k < NW; ++k) smem_add64(&rs[(ibase + k - 1) & Nmask], to_fx(W[k], fx_scale)); // rs is a shared-memory grid
}

struct SPArgs {
    const double *X, *V;
    double *v;
    double *xout;     // wrapped end-of-step position; during the sweeps it parks g1 = g(E1, X)
    const double *E;  // [3N]
    fx_t *rho;        // [3N]
    double *partials; // [2*gridDim.x]
    Ctrl *ctrl;
    long long P;
    double dt, fx_scale;
    int N, k;
};

template <int SHAPE>
__global__ void __launch_bounds__(PG_THREADS) sp_pass0(SPArgs a)
{
    using S = SpShape<SHAPE>;
    extern __shared__ double smem[];
    fx_t *r1 = reinterpret_cast<fx_t *>(smem);
    const int N = a.N, Nmask = N - 1;
    for (int n = threadIdx.x; n < N; n += blockDim.x) r1[n] = 0ULL;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.P; j += stride) {
        const double Xj = ld_stream(a.X + j);
        int ibase; double W[S::NW];
        S::weights((Xj + Xj) / 2, (double)N, ibase, W); // rho(X,X)
        sp_deposit(r1, ibase, W, a.fx_scale, Nmask);
    }
    __syncthreads();
    flush_grid(r1, a.rho, N);
}

// Dynamic shared memory: E1[N] doubles, r2[N], r3[N] fixed point.
template <int SHAPE>
__global__ void __launch_bounds__(PG_THREADS) sp_pass1(SPArgs a)
{
    using S = SpShape<SHAPE>;
    extern __shared__ double smem[];
    const int N = a.N, Nmask = N - 1;
    double *E1 = smem;
    fx_t *r2 = reinterpret_cast<fx_t *>(smem + N), *r3 = r2 + N;
    for (int n = threadIdx.x; n < N; n += blockDim.x) { E1[n] = a.E[n]; r2[n] = 0ULL; r3[n] = 0ULL; }
    __syncthreads();
    const double dN = (double)N, dt = a.dt;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.P; j += stride) {
        const double Xj = ld_stream(a.X + j), Vj = ld_stream(a.V + j);
        int ibase; double W[S::NW];
        S::weights((Xj + Xj) / 2, dN, ibase, W);
        st_stream(a.xout + j, sp_gather(E1, ibase, W, Nmask));    // g1, reused by every sweep (:12)
        const double xj = Xj + (Vj + Vj) / 2 * dt;                   // x_1 (v = V)
        S::weights((Xj + xj) / 2, dN, ibase, W);
        sp_deposit(r2, ibase, W, a.fx_scale, Nmask);       // rho(X,x)  :13
        S::weights((xj + xj) / 2, dN, ibase, W);
        sp_deposit(r3, ibase, W, a.fx_scale, Nmask);       // rho(x,x)  :15
    }
    __syncthreads();
    flush_grid(r2, a.rho + N, N);
    flush_grid(r3, a.rho + 2 * N, N);
}

// Dynamic shared memory: E2[N], E3[N] doubles, r2[N], r3[N] fixed point, 32 doubles scratch.
template <int SHAPE>
__global__ void __launch_bounds__(PG_THREADS) sp_passk(SPArgs a)
{
    using S = SpShape<SHAPE>;
    extern __shared__ double smem[];
    const int fk = a.ctrl->final_k;
    if (fk >= 0 && a.k > fk) return;
    const bool final = fk == a.k;
    const int N = a.N, Nmask = N - 1;
    double *E2 = smem, *E3 = smem + N, *scratch = smem + 4 * N;
    fx_t *r2 = reinterpret_cast<fx_t *>(smem + 2 * N), *r3 = r2 + N;
    for (int n = threadIdx.x; n < N; n += blockDim.x) { E2[n] = a.E[N + n]; E3[n] = a.E[2 * N + n]; r2[n] = 0ULL; r3[n] = 0ULL; }
    __syncthreads();
    const double dN = (double)N, dt = a.dt;
    double sv2 = 0.0, sv = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.P; j += stride) {
        const double Xj = ld_stream(a.X + j), Vj = ld_stream(a.V + j);
        double vj = a.k == 1 ? Vj : ld_stream(a.v + j);
        const double g1 = ld_stream(a.xout + j);
        double xj = Xj + (vj + Vj) / 2 * dt;                          // x_k from v_{k-1}  :11
        int ibase; double W[S::NW];
        S::weights((Xj + xj) / 2, dN, ibase, W);
        const double g2 = sp_gather(E2, ibase, W, Nmask);
        S::weights((xj + xj) / 2, dN, ibase, W);
        const double g3 = sp_gather(E3, ibase, W, Nmask);
        vj = Vj + g1 * dt / 6;                                        // :12
        vj = vj + g2 * (4 * dt) / 6;                                  // :14
        vj = vj + g3 * dt / 6;                                        // :16
        st_stream(a.v + j, vj);
        if (final) {
            st_stream(a.xout + j, jl_mod1(xj));                       // x.=mod.(x,1)  :17
            sv2 = fma(vj, vj, sv2);
            sv += vj;
            continue;
        }
        xj = Xj + (vj + Vj) / 2 * dt;                                 // x_{k+1}
        S::weights((Xj + xj) / 2, dN, ibase, W);
        sp_deposit(r2, ibase, W, a.fx_scale, Nmask);
        S::weights((xj + xj) / 2, dN, ibase, W);
        sp_deposit(r3, ibase, W, a.fx_scale, Nmask);
    }
    __syncthreads();
    if (final) {
        sv2 = block_sum(sv2, scratch);
        sv = block_sum(sv, scratch);
        if (threadIdx.x == 0) { a.partials[2 * blockIdx.x] = sv2; a.partials[2 * blockIdx.x + 1] = sv; }
    } else {
        flush_grid(r2, a.rho + N, N);
        flush_grid(r3, a.rho + 2 * N, N);
    }
}

// Solve of rows 2 and 3 (two blocks) + the Frobenius-norm isapprox(F, E) over the 3 x N matrix; row 1 (E1) does
// not change inside a step, so it only enters through its norm (ctrl->normE1sq, stored by the E1 solve).
struct SolveSPArgs {
    unsigned long long *rho_fx; // [3N]
    double *rho_last;           // [N]: keeps rho3 = rho(x,x), the script's `r` after a step
    double *E;                  // [3N]
    const double2 *tw;
    Ctrl *ctrl;
    double w, fx_inv, rtol, atol;
    int N, lg, k, max_sweeps;
};

__global__ void __launch_bounds__(1024) solve_simpson23_kernel(SolveSPArgs a)
{
    extern __shared__ double smem[];
    if (a.ctrl->final_k >= 0) return;
    const int N = a.N, row = 1 + blockIdx.x;
    double *re = smem, *im = smem + N, *scratch = smem + 2 * N;
    unsigned long long *rho = a.rho_fx + (size_t)row * N;
    double *E = a.E + (size_t)row * N;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double r = (double)(long long)rho[n] * a.fx_inv * a.w;
        rho[n] = 0ULL;
        if (row == 2) a.rho_last[n] = r;
        re[n] = r; im[n] = 0.0;
    }
    __syncthreads();
    fft_smem4<false>(re, im, N, 1, 1, 0, a.tw, N);
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
        int s = bitrev(p, a.lg);
        if (s == 0) { re[p] = 0.0; im[p] = 0.0; }
        else {
            double kk = (s <= N / 2) ? (double)s : (double)(s - N);
            double b = TWO_PI * kk;
            double zr = re[p], zi = im[p];
            re[p] = zi / b;
            im[p] = -zr / b;
        }
    }
    __syncthreads();
    fft_smem4<true>(re, im, N, 1, 1, 0, a.tw, N);
    double d2 = 0.0, f2 = 0.0, e2 = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double e = re[n] / (double)N, f = E[n];
        E[n] = e;
        double d = f - e;
        d2 = fma(d, d, d2); f2 = fma(f, f, f2); e2 = fma(e, e, e2);
    }
    d2 = block_sum(d2, scratch);
    f2 = block_sum(f2, scratch);
    e2 = block_sum(e2, scratch);
    if (threadIdx.x == 0) {
        Ctrl *c = a.ctrl;
        if (row == 2) c->sumE2 = e2; // sum(E[end,:].^2)  :18
        c->sp_acc[blockIdx.x * 3 + 0] = d2; c->sp_acc[blockIdx.x * 3 + 1] = f2; c->sp_acc[blockIdx.x * 3 + 2] = e2;
        __threadfence();
        if (atomicAdd(&c->sp_arrive, 1u) == 1u) { // second block to finish decides (fixed summation order: row 2 + row 3)
            __threadfence();
            volatile double *acc = c->sp_acc;
            double D2 = acc[0] + acc[3], F2 = c->normE1sq + (acc[1] + acc[4]), E2s = c->normE1sq + (acc[2] + acc[5]);
            double d = sqrt(D2), m = fmax(sqrt(F2), sqrt(E2s));
            bool conv = isfinite(d) && d <= fmax(a.atol, a.rtol * m);
            c->sweeps = a.k;
            c->sp_arrive = 0u;
            if (conv || a.k >= a.max_sweeps) c->final_k = a.k;
        }
    }
}

} // namespace pg
