// pg_peer.cuh -- sum of the per-GPU charge grids over NVLink peer memory, fused into the field solve.
//
// The reference sums its per-thread grids with `phi .= sum(ns, dims=3)` (src/Electrostatic2D3V.jl:139-141); across GPUs
// that sum was one ncclAllReduce per sweep, enqueued by the host for ALL max_sweeps sweeps of a step because the host
// never learns the sweep count -- at 8 GPUs ten 32-us collectives per step, seven of them summing zeros.  Here every
// rank copies its integer grid into a buffer the other ranks can read (cudaIpc), raises a flag, and the solve kernel of
// every rank waits for the flags and adds up the nranks grids itself with plain loads over NVLink (64-bit integer adds
// in rank order: bit-identical on all ranks and to the NCCL sum).  A predicated-off sweep costs nothing.
//
// Protocol (seq = number of sweeps published so far, counted ON THE DEVICE in this rank's own header -- the launches of a
// device-driven loop (picgolf.cu) are the same graph nodes every iteration, so no host-side counter can ride in their
// arguments; every rank executes the same sweeps, so the counters agree; slot = seq & 1):
//   publish(seq):  wait until every peer's `done` >= the sequence number this slot last carried (nobody still reads
//                  it); copy rho -> data[slot], clear rho; fence; ready = seq.
//   solve(seq):    wait until every peer's `ready` >= seq; rho[n] = sum_q peer[q].data[slot][n]; then done = seq.
// A rank publishes sweep s+1 only after its own solve(s), and solve(s) needs every peer's publish(s): no rank is ever
// more than one real sweep ahead, and with the `done` handshake two slots suffice.  A wait gives up only after ~5 minutes
// (a rank that stalls for seconds -- checkpoint I/O, a re-sort, module load -- is simply waited for, as NCCL would); on
// time-out the error flag is raised, the solve poisons E with NaN and ends the step, and picgolf_synchronize and the getters
// return PICGOLF_ERR_NCCL, so a dead peer can neither hang the GPU for ever nor produce a silently wrong field.
#pragma once
#include "pg_common.cuh"

namespace pg {

constexpr int PEER_MAX = 16;
constexpr long long PEER_TIMEOUT_CYCLES = 600000000000LL; // default: ~5 min at 1.9 GHz (PICGOLF_PEER_TIMEOUT_S overrides, picgolf.cu)

struct PeerPub {
    unsigned long long ready;       // highest sweep sequence number whose grid this rank has published
    unsigned long long done;        // highest sequence number this rank has finished reading from all peers
    unsigned long long last_pub[2]; // per slot: the sequence number of its content
    unsigned long long flushes;     // this rank's cumulative flush counter (polynomial mode), summed next to the grid
    unsigned long long seq;         // sweeps this rank has published (device-side sequence counter; never reset)
    unsigned long long pad_[10];    // header = 128 bytes; fx_t data[2][ncell] follows
};

struct PeerArgs {
    PeerPub *peer[PEER_MAX]; // peer[rank] is this rank's own buffer
    int nranks, rank;        // nranks <= 1: not in use
    long long ncell;
    int *error;
    const unsigned long long *flush_src; // this rank's flush counter, or NULL
    long long timeout_cycles;            // a wait gives up after this many SM cycles
};

__device__ __forceinline__ fx_t *peer_data(PeerPub *p, int slot, long long ncell) { return reinterpret_cast<fx_t *>(p + 1) + (size_t)slot * ncell; }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void peer_wait(const unsigned long long *flag, unsigned long long need, int *error, long long timeout_cycles)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < need) {
        if (clock64() - t0 > timeout_cycles) { *error = 1; return; }
        __nanosleep(64);
    }
}

// One block.  rho: this rank's integer charge grid (cleared here, like the solve does on a single GPU).
__global__ void __launch_bounds__(1024) peer_publish_kernel(PeerArgs p, fx_t *rho, const int *final_k, int fixedpoint)
{
    if (fixedpoint && *final_k >= 0) return; // step already converged: predicated no-op, like the solve
    PeerPub *me = p.peer[p.rank];
    const unsigned long long seq = me->seq + 1ULL; // only this kernel writes it, at its very end
    const int slot = (int)(seq & 1ULL);
    if (threadIdx.x < p.nranks) peer_wait(&p.peer[threadIdx.x]->done, me->last_pub[slot], p.error, p.timeout_cycles);
    __syncthreads();
    fx_t *dst = peer_data(me, slot, p.ncell);
    for (long long n = threadIdx.x; n < p.ncell; n += blockDim.x) { dst[n] = rho[n]; rho[n] = 0ULL; }
    if (threadIdx.x == 0) me->flushes = p.flush_src ? *p.flush_src : 0ULL;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { me->last_pub[slot] = seq; me->seq = seq; st_release_sys(&me->ready, seq); }
}

// In the solve kernel (one block): wait for all grids of this sweep, then peer_sum(n) for every cell, then peer_done().
// Returns the sequence number of the sweep being gathered (this rank's latest publish).
__device__ __forceinline__ unsigned long long peer_gather_begin(const PeerArgs &p)
{
    const unsigned long long seq = p.peer[p.rank]->seq;
    if (threadIdx.x < p.nranks) peer_wait(&p.peer[threadIdx.x]->ready, seq, p.error, p.timeout_cycles);
    __syncthreads();
    return seq;
}
__device__ __forceinline__ long long peer_sum(const PeerArgs &p, unsigned long long seq, long long n)
{
    const int slot = (int)(seq & 1ULL);
    unsigned long long s = 0ULL;
    for (int q = 0; q < p.nranks; ++q) s += ld_relaxed_sys(peer_data(p.peer[q], slot, p.ncell) + n);
    return (long long)s;
}
// The same sum for PT grid cells of one thread (n0, n0 + stride, ...), with the PT loads from a rank all in flight before the first is
// used: peer_sum() alone walks rank by rank with one dependent NVLink round trip each -- 64 in a row for 8 points on 8 GPUs, which cost
// the 8-GPU step 0.2 ms (3 solves); this way a solve waits for nranks round trips.
template <int PT>
__device__ __forceinline__ void peer_sum_many(const PeerArgs &p, unsigned long long seq, int n0, int stride, long long (&out)[PT])
{
    const int slot = (int)(seq & 1ULL);
    unsigned long long acc[PT];
#pragma unroll
    for (int i = 0; i < PT; ++i) acc[i] = 0ULL;
    for (int q = 0; q < p.nranks; ++q) {
        const fx_t *src = peer_data(p.peer[q], slot, p.ncell) + n0;
        unsigned long long v[PT];
#pragma unroll
        for (int i = 0; i < PT; ++i) v[i] = ld_relaxed_sys(src + (long long)i * stride);
#pragma unroll
        for (int i = 0; i < PT; ++i) acc[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < PT; ++i) out[i] = (long long)acc[i];
}
// Sum of the ranks' flush counters; the first warp calls it (after peer_gather_begin), lane 0 holds the result.
__device__ __forceinline__ unsigned long long peer_flush_sum_warp(const PeerArgs &p)
{
    const int lane = threadIdx.x & 31;
    unsigned long long s = lane < p.nranks ? ld_relaxed_sys(&p.peer[lane]->flushes) : 0ULL;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
__device__ __forceinline__ unsigned long long peer_flush_sum(const PeerArgs &p) // after peer_gather_begin
{
    unsigned long long s = 0ULL;
    for (int q = 0; q < p.nranks; ++q) s += ld_relaxed_sys(&p.peer[q]->flushes);
    return s;
}
__device__ __forceinline__ void peer_gather_end(const PeerArgs &p, unsigned long long seq) // after a __syncthreads() that follows the last peer_sum
{
    if (threadIdx.x == 0) st_release_sys(&p.peer[p.rank]->done, seq);
}

// Touches every peer's header once (picgolf_peer_connect): cudaIpcOpenMemHandle maps peer memory lazily, and the first
// access over NVLink must not land inside a timed sweep.
__global__ void peer_touch_kernel(PeerArgs p, unsigned long long *sink)
{
    unsigned long long s = 0ULL;
    for (int q = 0; q < p.nranks; ++q) s += ld_relaxed_sys(&p.peer[q]->ready) + ld_relaxed_sys(peer_data(p.peer[q], 1, p.ncell) + (p.ncell - 1));
    if (s == 0xFFFFFFFFFFFFFFFFULL) *sink = s;
}

} // namespace pg
