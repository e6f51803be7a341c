// pg_sort.cuh -- counting sort of the particle arrays by grid cell (1D) or tile (2D).
// The reference sorts too (sortparticles!, src/Electrostatic2D3V.jl:57-62,159-161; src/PIC2D3V.jl:210-215)
// for cache locality; here cell order is what lets a warp accumulate its deposit in a private
// window without atomics (pg_kernels_1d.cuh).  A permutation array `pid` remembers each particle's
// original slot so picgolf_get_particles returns the caller's order.
//   1. sort_hist_kernel    per-block shared histogram -> global bin counts
//   2. sort_scan_kernel    exclusive scan of the bins (one block)
//   3. sort_scatter_kernel per-block histogram, one global reservation per (block, bin), local ranks
//                          from shared-memory atomics, payload scatter
// Summation order inside a bin is not reproducible (atomic ranks); the deterministic mode does not
// use this path.
#pragma once
#include "pg_common.cuh"

namespace pg {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 4;  // particles per thread in a scatter tile (their whole payload is held in registers)
constexpr int SORT_MAX_ARR = 5;

struct SortArgs {
    const double *in[SORT_MAX_ARR];
    double *out[SORT_MAX_ARR];
    const unsigned int *pid_in; // NULL: identity
    unsigned int *pid_out;
    unsigned int *bin_count;  // [nbins]
    unsigned int *bin_cursor; // [nbins] running start offsets (== end offsets once the scatter is done)
    unsigned int *bin_start;  // [nbins] start offsets (kept for the tiled 2D kernel); may be NULL
    long long P;
    int narr, nbins;
    int mode;   // 0: 1D key = cell of in[0];  1: 2D key = tile of (in[0], in[1]) on the unit box;  2: the same on an Lx x Ly box
    int N, NY;  // grid
    int tshift; // 2D: log2(tile edge in cells)
    int vsplit; // 1D: 1 -> key = 2*cell + (v >= 0): each beam keeps its own bins, so a bin drifts as a whole (pg_kernels_poly.cuh)
    double kx, ky; // mode 2: cells per unit length NX/Lx, NY/Ly (PIC2D3V GridParameters: positions live in (0, Lx] x (0, Ly])
    int sublg;  // 1D with vsplit: log2 of the position sub-bins per cell (key = ((2*cell + sign) << sublg) + sub): particles of a
                // bin stay ordered by position, so the rows of 64 particles a warp of fp_pass_poly evaluates lie in ONE cell
                // except where a cell boundary cuts through the bin
};

// Bin of a particle from its first two payload values (1D: x and v; 2D: x and y).
__device__ __forceinline__ int sort_key_of(const SortArgs &a, double p0, double p1)
{
    if (a.mode == 0) {
        int c = (int)rint(p0 * (double)a.N); // the stencil centre Int(round(x*N)): a bin shares its window rows
        if (!a.vsplit) return c & (a.N - 1);
        const double d = p0 * (double)a.N - (double)c; // offset from the cell centre, [-1/2, 1/2]
        const int sub = min((1 << a.sublg) - 1, max(0, (int)((d + 0.5) * (double)(1 << a.sublg))));
        return (((c & (a.N - 1)) * 2 + (p1 >= 0.0 ? 1 : 0)) << a.sublg) + sub;
    }
    if (a.mode == 2) {
        int ex = ((int)ceil(p0 * a.kx) - 1) & (a.N - 1);
        int ey = ((int)ceil(p1 * a.ky) - 1) & (a.NY - 1);
        return (ey >> a.tshift) * max(1, a.N >> a.tshift) + (ex >> a.tshift);
    }
    int cx = ((int)ceil(p0 * (double)a.N) - 1) & (a.N - 1);
    int cy = ((int)ceil(p1 * (double)a.NY) - 1) & (a.NY - 1);
    return (cy >> a.tshift) * (a.N >> a.tshift) + (cx >> a.tshift);
}
__device__ __forceinline__ int sort_key(const SortArgs &a, long long j)
{
    return sort_key_of(a, a.in[0][j], (a.mode != 0 || a.vsplit) ? a.in[1][j] : 0.0);
}

__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(SortArgs a)
{
    extern __shared__ unsigned int sh[];
    for (int b = threadIdx.x; b < a.nbins; b += blockDim.x) sh[b] = 0u;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.P; j += stride) atomicAdd(&sh[sort_key(a, j)], 1u);
    __syncthreads();
    for (int b = threadIdx.x; b < a.nbins; b += blockDim.x) {
        unsigned int c = sh[b];
        if (c) atomicAdd(&a.bin_count[b], c);
    }
}

// Exclusive scan of the bin counts -> start offsets; one block of 1024 threads, any number of bins.  Warp w owns the contiguous
// run [w*per, (w+1)*per) and walks it in coalesced chunks of 32 bins (lane-consecutive loads, shuffle scan, running carry): first
// to total its run, then -- after the 32 run totals have been scanned -- to write the offsets and clear the counts for the next
// sort.  (Round 1 gave every THREAD a contiguous run: uncoalesced, 0.63 ms for the 2^18 bins of the polynomial mode; now 0.09 ms in the ncu launch list.)
__global__ void __launch_bounds__(1024) sort_scan_kernel(unsigned int *bin_count, unsigned int *bin_cursor,
                                                        unsigned int *bin_start, int nbins)
{
    __shared__ unsigned int wtot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int per = ((nbins + 31) / 32 + 31) & ~31; // bins per warp, a multiple of 32
    const int lo = wid * per, hi = min(lo + per, nbins);
    unsigned int s = 0;
    for (int b = lo + lane; b < hi; b += 32) s += bin_count[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) wtot[wid] = s;
    __syncthreads();
    if (wid == 0) { // exclusive scan of the 32 run totals
        const unsigned int t = wtot[lane];
        unsigned int inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        wtot[lane] = inc - t;
    }
    __syncthreads();
    unsigned int run = wtot[wid];
    for (int b0 = lo; b0 < hi; b0 += 32) {
        const int b = b0 + lane;
        const unsigned int c = b < hi ? bin_count[b] : 0u;
        unsigned int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (b < hi) {
            const unsigned int off = run + inc - c; // exclusive
            bin_cursor[b] = off;
            if (bin_start) bin_start[b] = off;
            bin_count[b] = 0u; // ready for the next sort
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// Dynamic shared memory: 2*nbins u32.  A block takes tiles of SORT_THREADS * SORT_ITEMS particles: every thread first
// requests the whole payload of its SORT_ITEMS particles (NARR doubles + the id each, all loads independent and in flight
// together -- the bin is computed from the loaded x, y / x, v, nothing is read twice), then the tile is ranked (lanes of a warp
// with the same key are grouped with match.any and reserve their slots with one shared-memory atomic per group: on the
// nearly sorted arrays of a re-sort a whole warp row has one key), one global reservation per (tile, bin) follows, and the
// payload is scattered.  Round 1's form (keys read first, payload re-read later in groups of four, one shared atomic per
// particle) moved 3.1 TB/s on a re-sort of the 2D arrays; it waited on four dependent memory round trips per tile.
template <int NARR>
__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(SortArgs a)
{
    static_assert(NARR >= 2, "the bin is computed from the first two payload arrays");
    extern __shared__ unsigned int sh[];
    unsigned int *cnt = sh, *base = sh + a.nbins;
    const long long tile = (long long)SORT_THREADS * SORT_ITEMS;
    const long long ntiles = (a.P + tile - 1) / tile;
    const int lane = threadIdx.x & 31;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long j0 = t * tile;
        double val[SORT_ITEMS][NARR];
        unsigned int id[SORT_ITEMS], rank[SORT_ITEMS];
        int key[SORT_ITEMS];
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            const long long j = j0 + (long long)i * SORT_THREADS + threadIdx.x;
            if (j < a.P) {
#pragma unroll
                for (int q = 0; q < NARR; ++q) val[i][q] = __ldcs(a.in[q] + j);
                id[i] = a.pid_in ? __ldcs(a.pid_in + j) : (unsigned int)j;
            }
        }
        for (int b = threadIdx.x; b < a.nbins; b += blockDim.x) cnt[b] = 0u;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            const long long j = j0 + (long long)i * SORT_THREADS + threadIdx.x;
            key[i] = j < a.P ? sort_key_of(a, val[i][0], val[i][1]) : -1;
            const int k = key[i] >= 0 ? key[i] : -1 - lane; // dead lanes: unique negative keys
            const unsigned int grp = __match_any_sync(0xffffffffu, k);
            const int leader = __ffs(grp) - 1;
            unsigned int r = 0;
            if (k >= 0 && lane == leader) r = atomicAdd(&cnt[k], (unsigned int)__popc(grp));
            r = __shfl_sync(0xffffffffu, r, leader);
            rank[i] = r + (unsigned int)__popc(grp & ((1u << lane) - 1u));
        }
        __syncthreads();
        for (int b = threadIdx.x; b < a.nbins; b += blockDim.x) {
            unsigned int c = cnt[b];
            if (c) base[b] = atomicAdd(&a.bin_cursor[b], c);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < SORT_ITEMS; ++i) {
            if (key[i] >= 0) {
                const long long d = (long long)base[key[i]] + rank[i];
#pragma unroll
                for (int q = 0; q < NARR; ++q) a.out[q][d] = val[i][q];
                a.pid_out[d] = id[i];
            }
        }
        __syncthreads();
    }
}

// Counting sort for many bins (polynomial mode: up to 65536 (cell, sign v, sub-cell) bins, too many for per-block
// shared-memory tables): lanes of a warp holding the same key are grouped with match.any, one global atomic per group.
// On the nearly sorted arrays of a re-sort a warp row holds one or two keys.
__global__ void __launch_bounds__(SORT_THREADS) sort_hist_match_kernel(SortArgs a)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nrow = (a.P + 31) >> 5;
    const int lane = threadIdx.x & 31;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrow; r += stride >> 5) {
        const long long j = (r << 5) + lane;
        const int key = j < a.P ? sort_key(a, j) : -1 - lane; // dead lanes: unique negative keys
        const unsigned int grp = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && lane == __ffs(grp) - 1) atomicAdd(&a.bin_count[key], (unsigned int)__popc(grp));
    }
}

template <int NARR>
__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_match_kernel(SortArgs a)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nrow = (a.P + 31) >> 5;
    const int lane = threadIdx.x & 31;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrow; r += stride >> 5) {
        const long long j = (r << 5) + lane;
        const bool live = j < a.P;
        const int key = live ? sort_key(a, j) : -1 - lane;
        double val[NARR];
        unsigned int id = 0;
        if (live) {
#pragma unroll
            for (int q = 0; q < NARR; ++q) val[q] = __ldcs(a.in[q] + j);
            id = a.pid_in ? __ldcs(a.pid_in + j) : (unsigned int)j;
        }
        const unsigned int grp = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(grp) - 1;
        unsigned int base = 0;
        if (live && lane == leader) base = atomicAdd(&a.bin_cursor[key], (unsigned int)__popc(grp));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (live) {
            const long long d = (long long)base + __popc(grp & ((1u << lane) - 1u));
#pragma unroll
            for (int q = 0; q < NARR; ++q) a.out[q][d] = val[q];
            a.pid_out[d] = id;
        }
    }
}

// Grid of the scatter: as many blocks as are resident (the kernel loops over its tiles), never more than there are tiles.
template <int NARR>
static inline int sort_scatter_grid(int sms, long long count, int nbins)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sort_scatter_kernel<NARR>, SORT_THREADS, (size_t)nbins * 8) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 2;
    }
    const long long tile = (long long)SORT_THREADS * SORT_ITEMS;
    const long long ntiles = (count + tile - 1) / tile;
    return (int)(ntiles < 1 ? 1 : (ntiles < (long long)sms * per_sm ? ntiles : (long long)sms * per_sm));
}

// out[pid[i]] = in[i]: back to the caller's particle order.
__global__ void unsort_kernel(const double *in, const unsigned int *pid, double *out, long long P)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += stride) out[pid[i]] = in[i];
}

} // namespace pg
