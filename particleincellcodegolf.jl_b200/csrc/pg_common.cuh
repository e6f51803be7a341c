// pg_common.cuh -- shared device/host helpers for libpicgolf (sm_100a).
// Compiled with -fmad=false: every a*b+c below rounds twice exactly like the Julia reference
// (Julia fuses only under @muladd); fused operations are written as explicit fma().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#define PG_THREADS 256
#define PG_WARP 32

namespace pg {

// ---- Julia Base semantics -------------------------------------------------------------

// Julia mod(a, 1) for Float64 (base/float.jl): r = rem(a,1); r==0 -> +0; r<0 -> r+1 (may round to
// exactly 1.0); else r.   src/NGPFourier.jl:2, src/GaussianFixedPoint.jl:9.
__device__ __forceinline__ double jl_mod1(double a)
{
    if (a > 0.0 && a < 1.0) return a; // nothing to wrap (nearly every call): the same bits as below without the FRND of trunc(), which
                                      // runs on the quarter-rate XU pipe -- 3 % of the TMA-staged NGP pass (same-box A/B, profiles/r2_t_*)
    double r = a - trunc(a); // == fmod(a, 1.0), exact
    if (r == 0.0) return 0.0;
    if (r < 0.0) return r + 1.0;
    return r;
}

// f(x) = Int(mod1(round(x*N), N)), src/NGPFourier.jl:3 -- returned 0-based (Julia cell - 1).
// round = ties-to-even (rint); mod1(0,N) = N.
__device__ __forceinline__ int ngp_cell0(double x, int N)
{
    long long r = __double2ll_rn(x * (double)N); // rint, ties to even
    long long m = r - 1;
    if (m == -1) return N - 1;                                 // mod1(0, N) = N
    if (m < 0 || m >= N) { m %= (long long)N; if (m < 0) m += N; } // x outside [0, 1]: never on the stepping path (x = mod(x, 1))
    return (int)m;
}

// Same for a power-of-two grid (all grids the library accepts): floored mod is a two's-complement mask,
// which spares the hot loops a 64-bit integer division.
__device__ __forceinline__ int ngp_cell0_pow2(double x, int N)
{
    long long r = __double2ll_rn(x * (double)N);
    return (int)((r - 1) & (long long)(N - 1));
}

// 0-based wrap of a Julia 1-based stencil index i (any sign): mod1(i,N)-1 == floormod(i-1, N).
__device__ __forceinline__ int wrap_cell0(int i, int N)
{
    int m = (i - 1) % N;
    return m < 0 ? m + N : m;
}

// unimod(x, n) = 0 < x <= n ? x : x > n ? x - n : x + n      src/Electrostatic2D3V.jl:83
__device__ __forceinline__ double unimod(double x, double n) { return (0.0 < x && x <= n) ? x : (x > n ? x - n : x + n); }
__device__ __forceinline__ int unimod(int x, int n) { return (0 < x && x <= n) ? x : (x > n ? x - n : x + n); }

// ---- fixed-point charge accumulation ------------------------------------------------------
// Deposits accumulate as round(weight * 2^frac) in 64-bit integers: integer addition is associative, so
// the grid is bit-identical whatever the order of the atomics (see pg_kernels_1d.cuh).
typedef unsigned long long fx_t;
__device__ __forceinline__ fx_t to_fx(double v, double fx_scale) { return (fx_t)__double2ll_rn(v * fx_scale); }
// The same rounding (to nearest, ties to even) for |v * fx_scale| < 2^51 without the conversion instruction (F2I.S64.F64 runs
// on the quarter-rate XU pipe): adding 1.5 * 2^52 leaves the rounded integer in the low mantissa bits.
__device__ __forceinline__ fx_t to_fx_small(double v, double fx_scale)
{
    const double M = 6755399441055744.0;
    return (fx_t)(__double_as_longlong(v * fx_scale + M) - __double_as_longlong(M));
}

// 64-bit integer add into SHARED memory as two native 32-bit atomics.  atomicAdd on a 64-bit shared word compiles to a compare-and-swap
// loop on sm_100a (ATOMS.CAST.SPIN.64); 32-bit adds are native (ATOMS.ADD).  The low-word add returns the old value, which tells whether
// THIS add carried; the carry rides along with the high-word add.  Every add propagates its own carry exactly once, so when all adds have
// landed the two words hold the same 64-bit sum (mod 2^64, two's complement included) the 64-bit atomic would have produced -- bit for bit,
// in any order.  (Readers must wait for a barrier, as with any shared-memory accumulation: between the two halves of an add the word is torn.)
// Measured on B200 (tools/f_rows_timing.py, 2^24 particles, N = 4096): the Simpson-1/3 passes, which deposit into two or three grids per
// particle, 8.5 -> 5.7 ms/step (Gaussian) and 3.1 -> 2.5 (area); the single-grid 1D2V pass is FASTER with the compare-and-swap form
// (0.72 vs 1.03 ms), so gauss_deposit_atomic keeps it.
__device__ __forceinline__ void smem_add64(fx_t *cell, fx_t v)
{
    unsigned int *w = reinterpret_cast<unsigned int *>(cell); // little endian: w[0] low, w[1] high
    const unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
    const unsigned int old = atomicAdd(w, lo);
    const unsigned int carry = old > ~lo ? 1u : 0u; // old + lo wrapped
    atomicAdd(w + 1, hi + carry); // unconditionally: skipping it when there is nothing to add costs more (a divergent branch) than it saves
}

// ---- reductions -------------------------------------------------------------------------

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in thread 0.  scratch: >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double *scratch)
{
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < nw ? scratch[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// ---- streaming loads/stores (particle arrays are touched once per pass) -------------------
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }

// FP64 pipe peak: 8 independent DFMA chains per thread (the denominator for the erf-shape kernels, which are
// bound by the FP64 pipe rather than by HBM; MEASURED_PEAKS.json has no fp64 figure).
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; ++i) {
        r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
        r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
    double s = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s; // never true; keeps the chains alive
}

// splitmix64: counter-based generator for the synthetic starts (NOT Julia's rand).
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double u01(uint64_t bits) { return (double)(bits >> 11) * (1.0 / 9007199254740992.0); }

} // namespace pg
