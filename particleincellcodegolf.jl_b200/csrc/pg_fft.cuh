// pg_fft.cuh -- shared-memory radix-2 FFT fused with the spectral Poisson solve.
//   1D: E = real.(ifft((xi = fft(rho)./ik; xi[1]*=0; xi))),  ik = 2pi*im*vcat(1, 1:N/2, -N/2+1:-1)
//       src/NGPFourier.jl:3,5; src/GaussianFixedPoint.jl:3,8  (FFTW conventions: forward
//       sum x[n] e^{-2 pi i nk/N} unnormalised, inverse 1/N).
//   2D: fft!(phi); phi[1,1]=0; Ex = phi*minvkk*kx; Ey = phi*minvkk*ky; ifft!(Ex); ifft!(Ey)
//       src/Electrostatic2D3V.jl:70-81,142-157.
// The forward transform is decimation-in-frequency (natural in, bit-reversed out), the pointwise
// divide runs on the bit-reversed spectrum, and the inverse is decimation-in-time (bit-reversed in,
// natural out), so no permutation pass is needed.  cuFFT is not used anywhere.
#pragma once
#include "pg_common.cuh"
#include "pg_peer.cuh"

namespace pg {

constexpr double TWO_PI = 6.283185307179586476925286766559; // Julia 2pi

// Device-resident loop control and per-step sums (one per handle).
struct Ctrl {
    int final_k;     // fixed point: sweep index whose pass finalises the step (-1: still iterating)
    int sweeps;      // solves executed in the current step
    int rows;        // diagnostics rows recorded so far
    int last_sweeps; // sweeps of the previous step
    long long step;  // steps completed
    double sumE2;    // sum(E.^2) of the latest solve (2D: sum(Ex^2+Ey^2))
    // Simpson-1/3 variant: ||E1||^2 of the step, per-row norm partials of the E2/E3 solve, arrival counter
    double normE1sq;
    double sp_acc[6];
    unsigned int sp_arrive;
    int fs_pred;     // fused re-sort (pg_kernels_poly.cuh): the passes k with k + 1 >= fs_pred count their bins; min of the last two steps' sweeps
    unsigned long long flush_global; // multi-GPU polynomial mode: sum over the ranks of the flush counters (pg_peer.cuh)
    unsigned long long flush_step;   // ... as the FIRST solve of a step sees it: everything up to the end of the previous step (the host's per-step probe)
    unsigned long long loop_sweeps;  // sweeps executed inside the device-driven loop since the particles were set (launch accounting)
    unsigned long long loop_fs_sweeps; // ... of which sweeps of a re-sorting step, which launch cp_fs_scan_kernel as well
};

// Sweep index of a launch.  The fixed schedule (stage timers on, NCCL reduction) passes it as a kernel argument; inside the
// device-driven loop (a CUDA-graph WHILE node, picgolf.cu) the same kernel nodes run every iteration, so k < 0 means "read it
// from Ctrl": the solve of an iteration is sweep ctrl->sweeps + 1 and records it, everything after it in the iteration reads
// ctrl->sweeps.
__device__ __forceinline__ int sweep_index(int k, const Ctrl *c) { return k >= 0 ? k : c->sweeps; }

// Batched in-place FFT on shared memory.  Element e of batch b lives at [b*bstride + e*estride].
// tw[k] = exp(-2 pi i k / twN), k < twN/2, n divides twN.  All threads of the block participate.
template <bool INVERSE>
__device__ __forceinline__ void fft_smem(double *re, double *im, int n, int batch, int estride, int bstride,
                                         const double2 *__restrict__ tw, int twN)
{
    const int nh = n >> 1;
    const int total = batch * nh;
    if (!INVERSE) {
        for (int half = nh; half >= 1; half >>= 1) {
            const int tws = twN / (2 * half);
            for (int t = threadIdx.x; t < total; t += blockDim.x) {
                int b = t / nh, u = t - b * nh;
                int pos = u & (half - 1), grp = u / half;
                int i = b * bstride + (grp * 2 * half + pos) * estride, j = i + half * estride;
                double2 w = __ldg(&tw[pos * tws]);
                double ar = re[i], ai = im[i], br = re[j], bi = im[j];
                re[i] = ar + br; im[i] = ai + bi;
                double dr = ar - br, di = ai - bi;
                re[j] = dr * w.x - di * w.y;
                im[j] = dr * w.y + di * w.x;
            }
            __syncthreads();
        }
    } else {
        for (int half = 1; half <= nh; half <<= 1) {
            const int tws = twN / (2 * half);
            for (int t = threadIdx.x; t < total; t += blockDim.x) {
                int b = t / nh, u = t - b * nh;
                int pos = u & (half - 1), grp = u / half;
                int i = b * bstride + (grp * 2 * half + pos) * estride, j = i + half * estride;
                double2 w = __ldg(&tw[pos * tws]); // conjugate below
                double br = re[j], bi = im[j];
                double tr = br * w.x + bi * w.y;
                double ti = bi * w.x - br * w.y;
                double ar = re[i], ai = im[i];
                re[j] = ar - tr; im[j] = ai - ti;
                re[i] = ar + tr; im[i] = ai + ti;
            }
            __syncthreads();
        }
    }
}

// Radix-2^2 form of fft_smem: two consecutive radix-2 stages are carried out in registers on four elements, with the SAME
// butterflies, twiddle-table entries and operation order as fft_smem -- the results are bit-identical -- but half the
// shared-memory round trips and block barriers (N = 4096: 6 + 6 passes instead of 12 + 12).  A leftover single stage
// (odd log2 n) runs first (forward) or last (inverse) as a plain radix-2 pass.
template <bool INVERSE>
__device__ __forceinline__ void fft_smem4(double *re, double *im, int n, int batch, int estride, int bstride,
                                          const double2 *__restrict__ tw, int twN)
{
#ifdef PG_FFT_RADIX2 // A/B builds (tools/fft_radix_check.py): plain radix-2 passes
    fft_smem<INVERSE>(re, im, n, batch, estride, bstride, tw, twN);
    return;
#endif
    const int nh = n >> 1, nq = n >> 2;
    int lg = 0;
    while ((1 << lg) < n) ++lg;
    auto radix2 = [&](int half) { // one stage of fft_smem
        const int tws = twN / (2 * half), total = batch * nh;
        for (int t = threadIdx.x; t < total; t += blockDim.x) {
            int b = t / nh, u = t - b * nh;
            int pos = u & (half - 1), grp = u / half;
            int i = b * bstride + (grp * 2 * half + pos) * estride, j = i + half * estride;
            double2 w = __ldg(&tw[pos * tws]);
            if (!INVERSE) {
                double ar = re[i], ai = im[i], br = re[j], bi = im[j];
                re[i] = ar + br; im[i] = ai + bi;
                double dr = ar - br, di = ai - bi;
                re[j] = dr * w.x - di * w.y;
                im[j] = dr * w.y + di * w.x;
            } else {
                double br = re[j], bi = im[j];
                double tr = br * w.x + bi * w.y;
                double ti = bi * w.x - br * w.y;
                double ar = re[i], ai = im[i];
                re[j] = ar - tr; im[j] = ai - ti;
                re[i] = ar + tr; im[i] = ai + ti;
            }
        }
        __syncthreads();
    };
    if (n < 4) { // n = 2: a single stage
        if (n == 2) radix2(1);
        return;
    }
    const int total = batch * nq;
    if (!INVERSE) {
        int half = nh;
        if (lg & 1) { radix2(half); half >>= 1; }
        for (; half >= 2; half >>= 2) { // stages `half` and `half/2`
            const int q = half >> 1;
            const int tws1 = twN / (2 * half), tws2 = twN / half;
            for (int t = threadIdx.x; t < total; t += blockDim.x) {
                int b = t / nq, u = t - b * nq;
                int pos = u & (q - 1), grp = u / q;
                int ia = b * bstride + (grp * 2 * half + pos) * estride;
                int ib = ia + q * estride, ic = ia + half * estride, id = ic + q * estride;
                double2 w1 = __ldg(&tw[pos * tws1]), w1q = __ldg(&tw[(pos + q) * tws1]), w2 = __ldg(&tw[pos * tws2]);
                double ar = re[ia], ai = im[ia], br = re[ib], bi = im[ib], cr = re[ic], ci = im[ic], dr = re[id], di = im[id];
                // stage `half`: (a,c) with w1, (b,d) with w1q
                double a1r = ar + cr, a1i = ai + ci, er = ar - cr, ei = ai - ci;
                double c1r = er * w1.x - ei * w1.y, c1i = er * w1.y + ei * w1.x;
                double b1r = br + dr, b1i = bi + di, fr = br - dr, fi = bi - di;
                double d1r = fr * w1q.x - fi * w1q.y, d1i = fr * w1q.y + fi * w1q.x;
                // stage `half/2`: (a',b') and (c',d') with w2
                re[ia] = a1r + b1r; im[ia] = a1i + b1i;
                double gr = a1r - b1r, gi = a1i - b1i;
                re[ib] = gr * w2.x - gi * w2.y; im[ib] = gr * w2.y + gi * w2.x;
                re[ic] = c1r + d1r; im[ic] = c1i + d1i;
                double hr = c1r - d1r, hi = c1i - d1i;
                re[id] = hr * w2.x - hi * w2.y; im[id] = hr * w2.y + hi * w2.x;
            }
            __syncthreads();
        }
    } else {
        int q = 1;
        for (int done = 0; done + 2 <= lg; done += 2, q <<= 2) { // stages `q` and `2q`
            const int half = 2 * q;
            const int tws1 = twN / (2 * q), tws2 = twN / (2 * half);
            for (int t = threadIdx.x; t < total; t += blockDim.x) {
                int b = t / nq, u = t - b * nq;
                int pos = u & (q - 1), grp = u / q;
                int ia = b * bstride + (grp * 4 * q + pos) * estride;
                int ib = ia + q * estride, ic = ia + half * estride, id = ic + q * estride;
                double2 w1 = __ldg(&tw[pos * tws1]), w2 = __ldg(&tw[pos * tws2]), w2q = __ldg(&tw[(pos + q) * tws2]);
                double ar = re[ia], ai = im[ia], br = re[ib], bi = im[ib], cr = re[ic], ci = im[ic], dr = re[id], di = im[id];
                // stage `q`: (a,b) and (c,d) with conj(w1)
                double tr = br * w1.x + bi * w1.y, ti = bi * w1.x - br * w1.y;
                double b1r = ar - tr, b1i = ai - ti, a1r = ar + tr, a1i = ai + ti;
                double ur = dr * w1.x + di * w1.y, ui = di * w1.x - dr * w1.y;
                double d1r = cr - ur, d1i = ci - ui, c1r = cr + ur, c1i = ci + ui;
                // stage `2q`: (a',c') with conj(w2), (b',d') with conj(w2q)
                double vr = c1r * w2.x + c1i * w2.y, vi = c1i * w2.x - c1r * w2.y;
                re[ic] = a1r - vr; im[ic] = a1i - vi; re[ia] = a1r + vr; im[ia] = a1i + vi;
                double xr = d1r * w2q.x + d1i * w2q.y, xi = d1i * w2q.x - d1r * w2q.y;
                re[id] = b1r - xr; im[id] = b1i - xi; re[ib] = b1r + xr; im[ib] = b1i + xi;
            }
            __syncthreads();
        }
        if (lg & 1) radix2(nh);
    }
}

// ---------------------------------------------------------------------------------------------
// Register-blocked Stockham FFT for the 1D solve (N >= 512, a power of two, N/8 threads): every thread holds 8 points, a pass is
// one radix-8 butterfly per thread (a closing radix-4 / radix-2 pass when log2 N is not a multiple of 3), so N = 4096 takes 4
// passes with 2 barriers each instead of the 6 radix-2^2 passes of fft_smem4, with 128-bit shared-memory accesses (interleaved
// re/im) and natural-order output (no bit reversal between the two transforms of a solve).  Reads of a pass are contiguous
// (x[i + r N/R]); its writes go to ((i - k) R + k) + r p, k = i mod p -- stride R on the first pass, which the padding
// (one 16-byte slot after every 8 points) spreads over all banks.  N = 8192: 16 points (two butterflies) per thread, so that the block
// stays at 512 threads and the butterflies in registers.  Forward transform only: the solve's inverse is
// real(ifft(xi)) = real(fft(conj xi)) / N.  tw[m] = exp(-2 pi i m / N), m < N/2.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int st_phys(int i) { return i + (i >> 3); }
__host__ __device__ inline size_t st_points(int N) { return (size_t)N + (size_t)(N >> 3); }

__device__ __forceinline__ double2 st_cmul(double2 a, double2 w) { return make_double2(fma(a.x, w.x, -(a.y * w.y)), fma(a.x, w.y, a.y * w.x)); }
__device__ __forceinline__ double2 st_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 st_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 st_mulmi(double2 a) { return make_double2(a.y, -a.x); } // a * (-i)
__device__ __forceinline__ double2 st_tw(const double2 *__restrict__ tw, int m, int N)
{
    const int h = N >> 1;
    double2 w = __ldg(&tw[m & (h - 1)]);
    if (m & h) { w.x = -w.x; w.y = -w.y; }
    return w;
}

// In-register DFT of R points, natural order in and out (e^{-i theta} convention).
template <int R>
__device__ __forceinline__ void st_dft(double2 (&u)[R])
{
    if constexpr (R == 2) {
        const double2 a = st_add(u[0], u[1]), b = st_sub(u[0], u[1]);
        u[0] = a; u[1] = b;
    } else if constexpr (R == 4) {
        const double2 p0 = st_add(u[0], u[2]), p1 = st_add(u[1], u[3]);
        const double2 q0 = st_sub(u[0], u[2]), q1 = st_mulmi(st_sub(u[1], u[3]));
        u[0] = st_add(p0, p1); u[2] = st_sub(p0, p1);
        u[1] = st_add(q0, q1); u[3] = st_sub(q0, q1);
    } else {
        constexpr double H = 0.70710678118654752440; // 1/sqrt(2)
        double2 s[4], t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[j] = st_add(u[j], u[j + 4]); t[j] = st_sub(u[j], u[j + 4]); }
        // t_j *= w8^j, w8 = (1 - i)/sqrt(2)
        t[1] = make_double2((t[1].x + t[1].y) * H, (t[1].y - t[1].x) * H);
        t[2] = st_mulmi(t[2]);
        t[3] = make_double2((t[3].y - t[3].x) * H, -(t[3].x + t[3].y) * H);
        st_dft<4>(s);
        st_dft<4>(t);
#pragma unroll
        for (int m = 0; m < 4; ++m) { u[2 * m] = s[m]; u[2 * m + 1] = t[m]; }
    }
}

// One pass: p = product of the radices of the passes before it.  blockDim.x == N/PT.
template <int R, int PT>
__device__ __forceinline__ void st_pass(double2 *buf, int N, int p, const double2 *__restrict__ tw)
{
    constexpr int M = PT / R; // butterflies per thread
    const int T = N / R;
    double2 u[M][R];
    int jb[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int i = threadIdx.x + m * blockDim.x;
        const int k = i & (p - 1);
        jb[m] = (i - k) * R + k;
        double2 w[R];
        if (p > 1) { // the table reads go first: they are the long-latency ones.  w^1, w^2, w^4 from the table, the others one product away
            const int tws = N / (p * R);
#pragma unroll
            for (int r = 1; r < R; r <<= 1) w[r] = st_tw(tw, k * r * tws, N);
            if constexpr (R >= 4) w[3] = st_cmul(w[1], w[2]);
            if constexpr (R == 8) { w[5] = st_cmul(w[1], w[4]); w[6] = st_cmul(w[2], w[4]); w[7] = st_cmul(w[3], w[4]); }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) u[m][r] = buf[st_phys(i + r * T)];
        if (p > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) u[m][r] = st_cmul(u[m][r], w[r]);
        }
        st_dft<R>(u[m]);
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int r = 0; r < R; ++r) buf[st_phys(jb[m] + r * p)] = u[m][r];
    __syncthreads();
}

template <int PT>
__device__ __forceinline__ void fft_stockham8(double2 *buf, int N, int lg, const double2 *__restrict__ tw)
{
    int p = 1;
    for (int done = 0; done + 3 <= lg; done += 3, p <<= 3) st_pass<8, PT>(buf, N, p, tw);
    if (lg % 3 == 2) st_pass<4, PT>(buf, N, p, tw);
    else if (lg % 3 == 1) st_pass<2, PT>(buf, N, p, tw);
}

// Launch shape of solve1d_kernel: N >= 512 runs the register-blocked transform (N/8 threads, padded interleaved points),
// smaller grids the radix-2^2 one (N/2 threads, separate planes).  32 doubles of reduction scratch follow the points.
__host__ __device__ inline bool solve1d_stockham(int N) { return N >= 512 && N <= 8192; }
__host__ __device__ inline int solve1d_points(int N) { return N > 4096 ? 16 : 8; } // grid points per thread of the register-blocked transform
__host__ __device__ inline int solve1d_threads(int N) { return solve1d_stockham(N) ? N / solve1d_points(N) : (N / 2 < 32 ? 32 : (N / 2 > 1024 ? 1024 : N / 2)); }
__host__ __device__ inline size_t solve1d_smem_bytes(int N) { return (solve1d_stockham(N) ? st_points(N) * 16 : (size_t)N * 16) + 32 * 8; }

__device__ __forceinline__ int bitrev(int v, int lg) { return (int)(__brev((unsigned)v) >> (32 - lg)); }

struct Solve1DArgs {
    const double *rho_in;           // stage entry only: fp64 charge density given by the caller
    unsigned long long *rho_fx;     // [N] integer deposit grid summed over ranks (NGP counts or Gaussian fixed
                                    // point): rho = fx * fx_inv * w; zeroed after it is read
    double *rho_last;               // [N] copy kept for picgolf_get_fields
    double *E;                      // [N] in: previous field (the reference's F), out: new field
    const double2 *tw;              // twiddles for size N
    Ctrl *ctrl;
    double w, fx_inv, rtol, atol;   // fx_inv = 2^-frac
    int N, lg, fixedpoint, k, max_sweeps;
    int store_normE1; // Simpson variant: this is the E1 solve of the step
    double *hist;     // 1D2V: time-averaged field history column Es[:,ti] += E (NGP1D2V.jl:57), or NULL
    PeerArgs peer;    // multi-GPU: sum the ranks' grids over peer memory here (pg_peer.cuh); nranks <= 1: rho_fx is the sum
    int flush_slot;   // NCCL path, polynomial mode: rho_fx[N] holds the all-reduced flush counter
    cudaGraphConditionalHandle cond; // device-driven loop: handle of the WHILE node this solve sits in (0: fixed schedule)
};

#ifdef PG_SOLVE_PROF // measurement builds only (tools/solve_prof.py): SM clock at the phase boundaries of the latest solve
__device__ long long g_solve_prof[8];
#define PG_PROF(i) do { __syncthreads(); if (threadIdx.x == 0) g_solve_prof[i] = clock64(); } while (0)
#else
#define PG_PROF(i) do { } while (0)
#endif

__device__ __forceinline__ void solve1d_decide(const Solve1DArgs &a, double d2, double f2, double e2, double *scratch, int k, bool peer_failed);

// One block.  Dynamic shared memory: 2*N doubles + 32.  (N >= 512: solve1d_stock_kernel below.)
__global__ void __launch_bounds__(1024) solve1d_kernel(Solve1DArgs a)
{
    extern __shared__ __align__(16) double smem[];
    double *re = smem, *im = smem + a.N, *scratch = smem + 2 * a.N;
    if (a.fixedpoint && a.ctrl->final_k >= 0) { // step already converged: predicated no-op
        if (a.cond && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0u);
        return;
    }
    const int N = a.N;
    PG_PROF(0);
    const int k = a.k >= 0 ? a.k : a.ctrl->sweeps + 1; // read by every thread before thread 0 records it below
    const bool peers = a.peer.nranks > 1 && !a.rho_in;
    unsigned long long pseq = 0ULL;
    if (peers) {
        pseq = peer_gather_begin(a.peer);
        if (threadIdx.x == 0) { const unsigned long long fsum = peer_flush_sum(a.peer); a.ctrl->flush_global = fsum; if (k == 1) a.ctrl->flush_step = fsum; }
    }
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double r;
        if (a.rho_in) r = a.rho_in[n];
        else if (peers) r = (double)peer_sum(a.peer, pseq, n) * a.fx_inv * a.w; // the publish kernel cleared rho_fx
        else { r = (double)(long long)a.rho_fx[n] * a.fx_inv * a.w; a.rho_fx[n] = 0ULL; }
        a.rho_last[n] = r;
        re[n] = r; im[n] = 0.0;
    }
    __syncthreads();
    if (peers) peer_gather_end(a.peer, pseq);
    const bool peer_failed = peers && *a.peer.error != 0; // a peer never published (pg_peer.cuh): poison the field, end the step
    if (a.flush_slot && threadIdx.x == 0) { a.ctrl->flush_global = a.rho_fx[N]; if (k == 1) a.ctrl->flush_step = a.rho_fx[N]; a.rho_fx[N] = 0ULL; }
    PG_PROF(1);
    fft_smem4<false>(re, im, N, 1, 1, 0, a.tw, N);
    PG_PROF(2);
    // xi = fft(rho)./ik ; xi[1] *= 0.   z/(i b) = (Im z)/b - i (Re z)/b,  b = 2pi*kk
    for (int p = threadIdx.x; p < N; p += blockDim.x) {
        int s = bitrev(p, a.lg); // frequency slot held at position p
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
    PG_PROF(3);
    fft_smem4<true>(re, im, N, 1, 1, 0, a.tw, N);
    PG_PROF(4);
    double d2 = 0.0, f2 = 0.0, e2 = 0.0;
    const double dN = (double)N;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double e = re[n] / dN;
        if (peer_failed) e = __longlong_as_double(0x7ff8000000000000LL);
        double f = a.E[n];
        a.E[n] = e;
        if (a.hist) a.hist[n] += e;
        double d = f - e;
        d2 = fma(d, d, d2); f2 = fma(f, f, f2); e2 = fma(e, e, e2);
    }
    PG_PROF(5);
    solve1d_decide(a, d2, f2, e2, scratch, k, peer_failed);
}

// Norms of the solve and the convergence decision of the fixed point (thread 0 writes Ctrl and steers the WHILE node).
__device__ __forceinline__ void solve1d_decide(const Solve1DArgs &a, double d2, double f2, double e2, double *scratch, int k, bool peer_failed)
{
    d2 = block_sum(d2, scratch);
    f2 = block_sum(f2, scratch);
    e2 = block_sum(e2, scratch);
    PG_PROF(6);
    if (threadIdx.x == 0) {
        a.ctrl->sumE2 = e2;
        if (a.store_normE1) a.ctrl->normE1sq = e2;
        if (a.fixedpoint) {
            // LinearAlgebra.isapprox(F,E;rtol,atol): norm(F-E) <= max(atol, rtol*max(norm(F),norm(E)))
            double d = sqrt(d2);
            double m = fmax(sqrt(f2), sqrt(e2));
            bool conv = isfinite(d) && d <= fmax(a.atol, a.rtol * m);
            const bool last = conv || k >= a.max_sweeps || peer_failed;
            a.ctrl->sweeps = k;
            if (a.cond) a.ctrl->loop_sweeps += 1ULL;
            if (last) a.ctrl->final_k = k;
            if (a.cond) cudaGraphSetConditional(a.cond, last ? 0u : 1u); // the rest of this iteration still runs (pass k finalises)
        }
    }
}

// The solve for N >= 512 (solve1d_stockham): the same steps as solve1d_kernel around the register-blocked transform; every thread owns
// PT grid points in every phase, with all its global loads issued before anything depends on them.  Dynamic shared memory: solve1d_smem_bytes(N).
template <int PT>
__global__ void __launch_bounds__(512) solve1d_stock_kernel(Solve1DArgs a)
{
    extern __shared__ __align__(16) double smem[];
    double2 *buf = reinterpret_cast<double2 *>(smem);
    double *scratch = smem + 2 * st_points(a.N);
    if (a.fixedpoint && a.ctrl->final_k >= 0) { // step already converged: predicated no-op
        if (a.cond && threadIdx.x == 0) cudaGraphSetConditional(a.cond, 0u);
        return;
    }
    const int N = a.N;
    PG_PROF(0);
    const int k = a.k >= 0 ? a.k : a.ctrl->sweeps + 1; // read by every thread before thread 0 records it below
    const bool peers = a.peer.nranks > 1 && !a.rho_in;
    unsigned long long pseq = 0ULL;
    if (peers) {
        pseq = peer_gather_begin(a.peer);
        if (threadIdx.x < 32) { // PEER_MAX <= 32 ranks: one lane each
            const unsigned long long fsum = peer_flush_sum_warp(a.peer);
            if (threadIdx.x == 0) { a.ctrl->flush_global = fsum; if (k == 1) a.ctrl->flush_step = fsum; }
        }
    }
    {
        double r[PT];
        if (a.rho_in) {
#pragma unroll
            for (int q = 0; q < PT; ++q) r[q] = a.rho_in[threadIdx.x + q * blockDim.x];
        } else if (peers) {
            long long raw[PT];
            peer_sum_many<PT>(a.peer, pseq, threadIdx.x, blockDim.x, raw); // the publish kernel cleared rho_fx
#pragma unroll
            for (int q = 0; q < PT; ++q) r[q] = (double)raw[q] * a.fx_inv * a.w;
        } else {
            unsigned long long raw[PT];
#pragma unroll
            for (int q = 0; q < PT; ++q) raw[q] = a.rho_fx[threadIdx.x + q * blockDim.x];
#pragma unroll
            for (int q = 0; q < PT; ++q) r[q] = (double)(long long)raw[q] * a.fx_inv * a.w;
        }
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            const int n = threadIdx.x + q * blockDim.x;
            if (!a.rho_in && !peers) a.rho_fx[n] = 0ULL;
            a.rho_last[n] = r[q];
            buf[st_phys(n)] = make_double2(r[q], 0.0);
        }
    }
    __syncthreads();
    if (peers) peer_gather_end(a.peer, pseq);
    const bool peer_failed = peers && *a.peer.error != 0; // a peer never published (pg_peer.cuh): poison the field, end the step
    if (a.flush_slot && threadIdx.x == 0) { a.ctrl->flush_global = a.rho_fx[N]; if (k == 1) a.ctrl->flush_step = a.rho_fx[N]; a.rho_fx[N] = 0ULL; }
    PG_PROF(1);
    fft_stockham8<PT>(buf, N, a.lg, a.tw);
    PG_PROF(2);
    // xi = fft(rho)./ik ; xi[1] *= 0, stored CONJUGATED: real(ifft(xi)) = real(fft(conj xi))/N.  z/(i b) = (Im z)/b - i (Re z)/b,  b = 2pi*kk
#pragma unroll
    for (int q = 0; q < PT; ++q) {
        const int s = threadIdx.x + q * blockDim.x;
        double2 z = buf[st_phys(s)];
        if (s == 0) z = make_double2(0.0, 0.0);
        else {
            const double kk = (s <= N / 2) ? (double)s : (double)(s - N);
            const double ib = 1.0 / (TWO_PI * kk); // one division per grid point
            z = make_double2(z.y * ib, z.x * ib);
        }
        buf[st_phys(s)] = z;
    }
    __syncthreads();
    PG_PROF(3);
    fft_stockham8<PT>(buf, N, a.lg, a.tw);
    PG_PROF(4);
    double d2 = 0.0, f2 = 0.0, e2 = 0.0;
    const double iN = 1.0 / (double)N; // exact: N is a power of two
    {
        double f[PT], hst[PT];
#pragma unroll
        for (int q = 0; q < PT; ++q) f[q] = a.E[threadIdx.x + q * blockDim.x];
        if (a.hist) {
#pragma unroll
            for (int q = 0; q < PT; ++q) hst[q] = a.hist[threadIdx.x + q * blockDim.x];
        }
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            const int n = threadIdx.x + q * blockDim.x;
            double e = buf[st_phys(n)].x * iN;
            if (peer_failed) e = __longlong_as_double(0x7ff8000000000000LL);
            a.E[n] = e;
            if (a.hist) a.hist[n] = hst[q] + e;
            const double d = f[q] - e;
            d2 = fma(d, d, d2); f2 = fma(f[q], f[q], f2); e2 = fma(e, e, e2);
        }
    }
    PG_PROF(5);
    solve1d_decide(a, d2, f2, e2, scratch, k, peer_failed);
}

// ---------------------------------------------------------------------------------------------
// Grids that are not a power of two (NGP leapfrog only: `fft` in src/NGPFourier.jl:3-5 takes any N, and its ik vector is
// well formed for every EVEN N).  The same solve as solve1d_kernel, as two direct O(N^2) transforms: N <= 8192, so
// 2 x 67 M complex multiply-adds at most, spread over all SMs -- one WARP per output element (the lanes split the sum and
// combine with a fixed shuffle tree), 32 outputs per block (the block's copy of the tables is what a block pays for).  Twiddles e^{2 pi i m/N}, m < N, come from a table the host
// computed in long double (make_dft_twiddles), indexed with the exact (k n) mod N, so every term is accurate to an ulp
// whatever N is.
//   pass 1 (solve1d_dft_fwd): xi_k = (sum_n rho_n e^{-2 pi i k n/N}) / ik_k, xi_0 = 0           -> spec[k]
//   pass 2 (solve1d_dft_inv): E_n = (1/N) Re sum_k xi_k e^{+2 pi i k n/N}; clears the deposit grid; sum(E.^2) by the last block
// Shared memory of both: N double2 twiddles + 2 N doubles (rho / spectrum planes) + reduction scratch.
// ---------------------------------------------------------------------------------------------
struct SolveDftArgs {
    const double *rho_in;       // stage entry only
    unsigned long long *rho_fx; // [N] integer deposit grid summed over ranks; zeroed by pass 2
    double *rho_last, *E;
    const double2 *tw;          // [N] (cos, sin)(2 pi m / N)
    double2 *spec;              // [N] xi
    double *part;               // [gridDim.x] partial sums of E^2
    unsigned int *arrive;       // [1] blocks of pass 2 that have finished (reset by the last one)
    Ctrl *ctrl;
    double w, fx_inv;
    int N;
};
constexpr int DFT_THREADS = 256, DFT_OUT = 32; // outputs per block: each of the 8 warps takes four, one after the other
__host__ __device__ inline size_t dft_smem_bytes(int N) { return (size_t)N * (8 + 8 + 16) + 256; }
__host__ __device__ inline int dft_blocks(int N) { return (N + DFT_OUT - 1) / DFT_OUT; }

__global__ void __launch_bounds__(DFT_THREADS) solve1d_dft_fwd(SolveDftArgs a)
{
    extern __shared__ __align__(16) unsigned char dft_raw[];
    double2 *tw = reinterpret_cast<double2 *>(dft_raw);
    double *rho = reinterpret_cast<double *>(tw + a.N);
    const int N = a.N, lane = threadIdx.x & 31;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        tw[n] = a.tw[n];
        const double r = a.rho_in ? a.rho_in[n] : (double)(long long)a.rho_fx[n] * a.fx_inv * a.w;
        rho[n] = r;
        if (blockIdx.x == 0) a.rho_last[n] = r;
    }
    __syncthreads();
    for (int o = threadIdx.x >> 5; o < DFT_OUT; o += DFT_THREADS / 32) {
        const int k = blockIdx.x * DFT_OUT + o;
        if (k >= N) return;
        double sr = 0.0, si = 0.0;
        int idx = (int)(((long long)k * lane) % N);          // (k n) mod N for n = lane, lane + 32, ...
        const int step = (int)(((long long)k * 32) % N);
        for (int n = lane; n < N; n += 32) {
            const double2 t = tw[idx]; // e^{-i theta} = (cos, -sin)
            sr = fma(rho[n], t.x, sr);
            si = fma(-rho[n], t.y, si);
            idx += step; if (idx >= N) idx -= N;
        }
        sr = warp_sum(sr); si = warp_sum(si);
        if (lane == 0) {
            // xi = fft(rho)./ik ; xi[1] *= 0.   z/(i b) = (Im z)/b - i (Re z)/b,  b = 2pi*kk,  kk = vcat(1, 1:N/2, -N/2+1:-1)[k+1]
            double2 xi = make_double2(0.0, 0.0);
            if (k > 0) {
                const double kk = (k <= N / 2) ? (double)k : (double)(k - N);
                const double b = TWO_PI * kk;
                xi = make_double2(si / b, -sr / b);
            }
            a.spec[k] = xi;
        }
    }
}

__global__ void __launch_bounds__(DFT_THREADS) solve1d_dft_inv(SolveDftArgs a)
{
    extern __shared__ __align__(16) unsigned char dft_raw[];
    double2 *tw = reinterpret_cast<double2 *>(dft_raw);
    double *xr = reinterpret_cast<double *>(tw + a.N), *xi = xr + a.N;
    double *scratch = xi + a.N;
    __shared__ bool last;
    const int N = a.N, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < N; k += blockDim.x) { tw[k] = a.tw[k]; const double2 z = a.spec[k]; xr[k] = z.x; xi[k] = z.y; }
    __syncthreads();
    double e2 = 0.0;
    for (int o = threadIdx.x >> 5; o < DFT_OUT; o += DFT_THREADS / 32) {
        const int n = blockIdx.x * DFT_OUT + o;
        if (n >= N) break;
        double s = 0.0, u = 0.0; // two independent chains
        int idx = (int)(((long long)n * lane) % N);      // (k n) mod N for k = lane, lane + 32, ...
        const int step = (int)(((long long)n * 32) % N);
        for (int k = lane; k < N; k += 32) {
            const double2 t = tw[idx];
            s = fma(xr[k], t.x, s);  // Re (xr + i xi)(cos + i sin)
            u = fma(-xi[k], t.y, u);
            idx += step; if (idx >= N) idx -= N;
        }
        s = warp_sum(s + u);
        if (lane == 0) {
            const double e = s / (double)N;
            a.E[n] = e;
            if (!a.rho_in) a.rho_fx[n] = 0ULL;
            e2 += e * e;
        }
    }
    e2 = block_sum(e2, scratch);
    if (threadIdx.x == 0) {
        a.part[blockIdx.x] = e2;
        __threadfence();
        last = atomicAdd(a.arrive, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) { // fixed summation order: the same bits whichever block finishes last
        __threadfence();
        double t = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) t += a.part[b];
        a.ctrl->sumE2 = t;
        *a.arrive = 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// 2D solve: three kernels (x-rows forward, y-columns forward+invert+inverse, x-rows inverse).
// Z holds the complex spectrum with x in bit-reversed order after pass A.
// real(ifft(S)) only sees the Hermitian part of S; for real rho that makes Ex on the kx-Nyquist row
// and Ey on the ky-Nyquist column vanish, after which both spectra are Hermitian and ONE inverse
// transform of Ex^ + i Ey^ returns Ex in the real part and Ey in the imaginary part.
// ---------------------------------------------------------------------------------------------
struct Solve2DArgs {
    const double *rho_in; // stage entry only: fp64 charge density given by the caller
    fx_t *rho_fx;       // [NX*NY] column-major fixed-point deposit grid; rho = fx*fx_inv*w; zeroed after read
    double *rho_last;
    double2 *Z;         // [NX*NY] complex scratch
    double2 *E2;        // [NX*NY] (real(Ex), real(Ey)) per cell
    const double2 *twx, *twy;
    double *partials;   // per-block partial sums of Ex^2+Ey^2 (pass C)
    double w, fx_inv;
    int NX, NY, lgx, lgy;
};

constexpr int ROWS_PER_BLOCK = 4;
constexpr int COLS_PER_BLOCK = 8;

// Pass A: forward along x for ROWS_PER_BLOCK rows.  smem: 2*ROWS*NX doubles.
__global__ void __launch_bounds__(512) solve2d_rows_fwd(Solve2DArgs a)
{
    extern __shared__ double smem[];
    const int NX = a.NX, R = ROWS_PER_BLOCK;
    double *re = smem, *im = smem + R * NX;
    const int j0 = blockIdx.x * R;
    for (int t = threadIdx.x; t < R * NX; t += blockDim.x) {
        int r = t / NX, i = t - r * NX;
        size_t g = (size_t)i + (size_t)(j0 + r) * NX;
        double v;
        if (a.rho_in) v = a.rho_in[g];
        else { v = (double)(long long)a.rho_fx[g] * a.fx_inv * a.w; a.rho_fx[g] = 0ULL; }
        a.rho_last[g] = v;
        re[t] = v; im[t] = 0.0;
    }
    __syncthreads();
    fft_smem4<false>(re, im, NX, R, 1, NX, a.twx, NX);
    for (int t = threadIdx.x; t < R * NX; t += blockDim.x) {
        int r = t / NX, i = t - r * NX;
        a.Z[(size_t)i + (size_t)(j0 + r) * NX] = make_double2(re[t], im[t]);
    }
}

// Pass B: for COLS_PER_BLOCK x-positions: forward along y, multiply, inverse along y.
// smem layout [c][j] with row stride NY+1 to stagger banks.
__global__ void __launch_bounds__(512) solve2d_cols(Solve2DArgs a)
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
        double kx = TWO_PI * (double)(ix < NX / 2 ? ix : ix - NX);
        double ky = TWO_PI * (double)(iy < NY / 2 ? iy : iy - NY);
        double ar = re[c * LD + q], ai = im[c * LD + q];
        double zr = 0.0, zi = 0.0;
        if (ix != 0 || iy != 0) {
            double m = -1.0 / (kx * kx + ky * ky); // minvkk = (0, m)
            double tr = -(ai * m), ti = ar * m;    // phi*minvkk
            double exr = tr * kx, exi = ti * kx, eyr = tr * ky, eyi = ti * ky;
            if (ix == NX / 2) { exr = 0.0; exi = 0.0; }
            if (iy == NY / 2) { eyr = 0.0; eyi = 0.0; }
            zr = exr - eyi; zi = exi + eyr; // Ex^ + i Ey^
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

// Pass C: inverse along x; E2 = (re, im)/(NX*NY); per-block partial of sum(Ex^2+Ey^2).
__global__ void __launch_bounds__(512) solve2d_rows_inv(Solve2DArgs a)
{
    extern __shared__ double smem[];
    const int NX = a.NX, R = ROWS_PER_BLOCK;
    double *re = smem, *im = smem + R * NX, *scratch = smem + 2 * R * NX;
    const int j0 = blockIdx.x * R;
    for (int t = threadIdx.x; t < R * NX; t += blockDim.x) {
        int r = t / NX, i = t - r * NX;
        double2 z = a.Z[(size_t)i + (size_t)(j0 + r) * NX];
        re[t] = z.x; im[t] = z.y;
    }
    __syncthreads();
    fft_smem4<true>(re, im, NX, R, 1, NX, a.twx, NX);
    const double inv = (double)NX * (double)a.NY;
    double e2 = 0.0;
    for (int t = threadIdx.x; t < R * NX; t += blockDim.x) {
        int r = t / NX, i = t - r * NX;
        double ex = re[t] / inv, ey = im[t] / inv;
        a.E2[(size_t)i + (size_t)(j0 + r) * NX] = make_double2(ex, ey);
        e2 += ex * ex + ey * ey;
    }
    e2 = block_sum(e2, scratch);
    if (threadIdx.x == 0) a.partials[blockIdx.x] = e2;
}

} // namespace pg
