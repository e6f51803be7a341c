// picgolf.cu -- C ABI of libpicgolf.so (include/picgolf.h).  Host-side orchestration only: every
// number is produced by the CUDA kernels in pg_*.cuh; there is no CPU compute path and no fallback.
#include "../../include/picgolf.h"

#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <map>
#include <vector>

#include "pg_common.cuh"
#include "pg_fft.cuh"
#include "pg_gauss.cuh"
#include "pg_kernels_1d.cuh"
#include "pg_kernels_2d.cuh"
#include "pg_kernels_poly.cuh"
#include "pg_sort.cuh"
#include "pg_sort_policy.h"
#include "pg_kernels_simpson.cuh"

using namespace pg;

#ifndef PG_POLY_SUBLG
#define PG_POLY_SUBLG (CP_SUBLG + 2)
#endif
constexpr int POLY_MAXBINS = 1 << (16 + CP_SUBLG);
constexpr int POLY_SUBLG = PG_POLY_SUBLG; // polynomial mode: up to four position sub-bins per polynomial interval in every (cell, sign v) bin

#define PG_API extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define PG_CUDA(call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(PICGOLF_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

#define PG_TRY(expr)            \
    do {                        \
        int rc_ = (expr);       \
        if (rc_ != 0) return rc_; \
    } while (0)

static bool is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }
static int ilog2(int64_t n) { int l = 0; while (((int64_t)1 << l) < n) ++l; return l; }

// ------------------------------------------------------------------------------------------
// NCCL, resolved at run time (the process's already-loaded libnccl.so.2 if there is one)
// ------------------------------------------------------------------------------------------
namespace nccl {
typedef struct { char internal[128]; } UniqueId;
typedef void *Comm;
enum { Sum = 0 };
enum { Int64 = 4, Uint64 = 5, Float64 = 8 };
static void *lib = nullptr;
static int (*GetUniqueId)(UniqueId *) = nullptr;
static int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
static int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
static int (*CommDestroy)(Comm) = nullptr;
static const char *(*GetErrorString)(int) = nullptr;

static int load()
{
    if (AllReduce) return 0;
    const char *env = getenv("PICGOLF_NCCL_LIB");
    if (env && *env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // share the instance torch loaded
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(PICGOLF_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    GetUniqueId = (int (*)(UniqueId *))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (int (*)(Comm *, int, UniqueId, int))dlsym(lib, "ncclCommInitRank");
    AllReduce = (int (*)(const void *, void *, size_t, int, int, Comm, cudaStream_t))dlsym(lib, "ncclAllReduce");
    CommDestroy = (int (*)(Comm))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) {
        AllReduce = nullptr;
        return fail(PICGOLF_ERR_NCCL, "libnccl is missing required symbols");
    }
    return 0;
}
static int check(int rc, const char *what)
{
    if (rc == 0) return 0;
    return fail(PICGOLF_ERR_NCCL, "%s failed: %s", what, GetErrorString ? GetErrorString(rc) : "?");
}
} // namespace nccl

// ------------------------------------------------------------------------------------------
// stage timers: CUDA events recorded on the handle's stream, resolved lazily (never inside a step)
// ------------------------------------------------------------------------------------------
struct StageTimer {
    struct Span { int stage; cudaEvent_t a, b; bool done; };
    std::vector<Span> open;
    std::vector<cudaEvent_t> pool;
    double ms[5] = {0, 0, 0, 0, 0};
    bool enabled = false;
    cudaEvent_t get()
    {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    // begin() returns a span id for end(); spans may nest (ST_TOTAL encloses the stage spans).
    int begin(int stage, cudaStream_t s)
    {
        if (!enabled) return -1;
        Span sp{stage, get(), get(), false};
        cudaEventRecord(sp.a, s);
        open.push_back(sp);
        return (int)open.size() - 1;
    }
    void end(int id, cudaStream_t s)
    {
        if (id < 0 || id >= (int)open.size()) return;
        cudaEventRecord(open[id].b, s);
        open[id].done = true;
    }
    void drain()
    {
        for (auto &sp : open) {
            float t = 0.f;
            if (sp.done && cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess)
                ms[sp.stage] += t;
            else
                cudaGetLastError();
            pool.push_back(sp.a); pool.push_back(sp.b);
        }
        open.clear();
    }
    void destroy()
    {
        drain();
        for (auto e : pool) cudaEventDestroy(e);
        pool.clear();
    }
};
enum { ST_PARTICLES = 0, ST_REDUCE = 1, ST_SOLVE = 2, ST_SORT = 3, ST_TOTAL = 4 };

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct picgolf_handle_s {
    picgolf_config cfg;
    int device = 0, sms = 0;
    cudaStream_t stream = nullptr;
    int64_t first = 0, count = 0;
    bool is2d = false, fixedpoint = false, ngp = false, simpson = false, b1d2v = false, have_deposit = false;
    double *vy1 = nullptr, *hist = nullptr; // 1D2V: vy array, field history [N x T]
    double *snap[3] = {nullptr, nullptr, nullptr}; // 2D3V with cfg.field_history: Exs, Eys, phis [NX*NY x T]   Electrostatic2D3V.jl:171-173
    size_t smem_b1 = 0;
    size_t smem_sp1 = 0, smem_spk = 0;
    // particles (1D: xb/vb ping-pong for the fixed point; leapfrog and 2D use index 0)
    double *xb[2] = {nullptr, nullptr}, *vb[2] = {nullptr, nullptr};
    double *p2[2][5] = {{nullptr, nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr, nullptr}}; // 2D: x,y,vx,vy,vz ping-pong
    int par = 0;
    // grids
    double *rho_last = nullptr, *E = nullptr;
    unsigned long long *rho_fx = nullptr;                     // integer deposit grid (NGP counts / fixed point), 1D and 2D
    // Gaussian fixed point: the final pass of a step deposits the NEXT step's first charge into rho_next, not into
    // rho_fx, so that the all-reduces the host enqueues for the remaining (predicated-off) sweeps of this step only
    // ever sum zeros; the two grids swap roles at the end of every step.
    unsigned long long *rho_next = nullptr, *rho_base[2] = {nullptr, nullptr};
    double fx_scale = 1.0, fx_inv = 1.0;                      // 2^frac, 2^-frac
    double2 *tw = nullptr, *twy = nullptr, *Z = nullptr, *E2 = nullptr;
    double *epartials = nullptr;
    int64_t ncell = 0;
    int grid_rows = 1;
    // control / diagnostics
    Ctrl *ctrl = nullptr;
    double *partials = nullptr, *raw = nullptr;
    int64_t T = 1;
    int nblocks = 1, npart = 2;
    int pass_blocks = 0; // blocks of the particle pass that wrote this step's partial sums (0: nblocks)
    size_t smem_pass = 0, smem_lf = 0;
    bool ngp_tma = false;
    bool dft = false;            // grid is not a power of two (NGP leapfrog): solve1d_dft_fwd / solve1d_dft_inv instead of solve1d_kernel
    double2 *dft_spec = nullptr, *dft_tw = nullptr; double *dft_part = nullptr; unsigned int *dft_arrive = nullptr;
    int k2d = 0, k2d_a = 0, k2d_b = 0, k2d_c = 0; // 2D tile-sorted particle kernel: 0 = particles_2d3v_tiled, 2 = particles_2d3v_stream<a, b, c>
    size_t smem_ring = 0;
    bool have_particles = false;
    int64_t steps = 0, launches = 0;
    nccl::Comm comm = nullptr;
    int nranks = 1, rank = 0;
    // peer-memory reduction of the charge grid (pg_peer.cuh): this rank's exported buffer, the opened peer buffers,
    // the sweep sequence counter, and the device-side time-out flag
    PeerPub *peer_mine = nullptr, *peer_ptr[PEER_MAX] = {};
    bool peer_ok = false, peer_this_solve = false;
    int *peer_err = nullptr;
    StageTimer timer;
    // cell-sorted mode
    bool sorted = false, pid_valid = false, use_sorted_now = false;
    unsigned int *pid[2] = {nullptr, nullptr};
    int pidpar = 0;
    unsigned int *bin_count = nullptr, *bin_cursor = nullptr, *bin_start = nullptr, *item_off = nullptr;
    double fxw_scale = 1.0; int fx_shift = 0; // 2D tiled: shared-memory window format
    unsigned long long *slow_count = nullptr;
    int nbins = 0, K = 1, sort_every = 1, nblocks_sorted = 1;
    int64_t since_sort = 0, sorts = 0;
    size_t smem_sorted = 0;
    // cell-polynomial mode (pg_kernels_poly.cuh): per-cell gather polynomials and fixed-point moment grid
    bool poly = false;
    bool det = false; // deterministic = 1 on the polynomial path: integer moment / diagnostics sums (fp_pass_poly<., true>)
    double *Gpoly = nullptr;
    unsigned long long *Mg = nullptr;
    int nblocks_poly = 1, sublg = 0;
    size_t smem_poly = 0;
    // re-sort fused into the passes of a step (pg_kernels_poly.cuh): third velocity buffer (the final pass writes v to its slot
    // there while the work buffer is read in place; the two swap roles afterwards), whether this step is such a step
    double *vspare = nullptr;
    unsigned long long *fs_sync = nullptr; // (epoch, total) words of the multi-block bin scan
    bool fs_enabled = false, fs_now = false;
    int64_t fused_sorts = 0;
    int probe_age[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // steps since the last sort at the start of the probed step
    int grow_hold = 0;
    // adaptive re-sort interval (cfg.sort_every == 0): the slow-path counter is copied to pinned memory at every
    // sort and looked at, without synchronising, at the next one
    bool sort_auto = false;
    unsigned long long *slow_host = nullptr, slow_seen = 0;
    cudaEvent_t slow_ev = nullptr;
    bool slow_pending = false;
    bool force_sort = false, poly_quiet = false, probe_have_prev = false; // polynomial mode: per-step flush probe (probe_poly_flushes)
    int64_t probe_step[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    cudaEvent_t run_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // end-of-step markers
    int64_t steps_at_probe = 0, steps_at_probe_prev_steps = 0;
    // Simpson-1/3 schemes: CUDA graphs of one fixed-schedule step, one per ping-pong parity
    cudaGraphExec_t step_graph[4] = {nullptr, nullptr, nullptr, nullptr}; // [par + 2*have_deposit]
    int64_t graph_launches[4] = {0, 0, 0, 0};
    bool graph_failed = false;
    // Gaussian fixed point: the device-driven sweep loop.  One graph per (X/V parity, charge-grid parity, kernel family):
    //   WHILE (not converged) { [moments -> rho] [publish] solve [gather polynomials] particle pass }  ->  step_end
    // The solve kernel sets the WHILE condition (cudaGraphSetConditional), so exactly S sweeps are launched -- no
    // predicated-off launches, no host involvement (for _ in 0:9 ... && break, GaussianFixedPoint.jl:7).
    // keyed by everything a captured launch bakes in: the particle buffers in their current roles, the charge grid, the kernel family
    std::map<std::array<uintptr_t, 8>, cudaGraphExec_t> loop_graph;
    cudaStream_t cap_stream = nullptr;
    bool loop_failed = false, loop_off = false;
    // picgolf_step_streamed: a ring of three particle buffer sets (set 0 = the handle's own arrays) and two copy streams,
    // so that the upload of call n+1, the step of call n and the download of call n-1 overlap (PCIe is full duplex)
    struct BufSet { double *a[10]; };  // 1D: xb[0], xb[1], vb[0], vb[1];  2D: p2[0][0..4]
    BufSet sset[3] = {};
    int sset_n = 0, sset_cur = 0;      // arrays per set (0: ring not built yet); set currently installed in xb/vb/p2
    bool sset_used[3] = {false, false, false};
    cudaStream_t up_stream = nullptr, down_stream = nullptr;
    cudaEvent_t ev_up[3] = {nullptr, nullptr, nullptr}, ev_comp[3] = {nullptr, nullptr, nullptr}, ev_down[3] = {nullptr, nullptr, nullptr};
    int64_t stream_calls = 0;
    int loop_launches_per_sweep = 0;
    int64_t loop_steps = 0;
};

static int use_device(picgolf_handle h) { PG_CUDA(cudaSetDevice(h->device)); return 0; }

// After a synchronisation: did a peer-memory wait time out (pg_peer.cuh)?  The fields are NaN-poisoned in that case.
static int check_peer(picgolf_handle h)
{
    if (!h->peer_err) return 0;
    int err = 0;
    PG_CUDA(cudaMemcpy(&err, h->peer_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err) return fail(PICGOLF_ERR_NCCL, "peer-memory reduction: a rank never published its charge grid (timed out); the fields of this run are poisoned with NaN");
    return 0;
}


template <typename T>
static int dalloc(T **p, size_t n)
{
    PG_CUDA(cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}

static int make_twiddles(double2 **dst, int n)
{
    std::vector<double2> t(std::max(n / 2, 1));
    for (int k = 0; k < n / 2; ++k) {
        long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        t[k] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    PG_TRY(dalloc(dst, t.size()));
    PG_CUDA(cudaMemcpy(*dst, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice));
    return 0;
}

// (cos, sin)(2 pi m / n), m < n: the full circle for the direct transforms of grids that are not a power of two
static int make_dft_twiddles(double2 **dst, int n)
{
    std::vector<double2> t((size_t)std::max(n, 1));
    for (int m = 0; m < n; ++m) {
        long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)n;
        t[m] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    PG_TRY(dalloc(dst, t.size()));
    PG_CUDA(cudaMemcpy(*dst, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice));
    return 0;
}

template <typename K>
static int set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) PG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

// N >= 512: register-blocked Stockham transform (8 or 16 grid points per thread), else the radix-2^2 one.
static void launch_solve1d_kernel(const Solve1DArgs &a, cudaStream_t stream)
{
    const int N = a.N;
    if (!solve1d_stockham(N)) solve1d_kernel<<<1, solve1d_threads(N), solve1d_smem_bytes(N), stream>>>(a);
    else if (solve1d_points(N) == 16) solve1d_stock_kernel<16><<<1, solve1d_threads(N), solve1d_smem_bytes(N), stream>>>(a);
    else solve1d_stock_kernel<8><<<1, solve1d_threads(N), solve1d_smem_bytes(N), stream>>>(a);
}
static int set_smem_solve1d(int N)
{
    const size_t smem = solve1d_smem_bytes(N);
    if (!solve1d_stockham(N)) return set_smem(solve1d_kernel, smem);
    return solve1d_points(N) == 16 ? set_smem(solve1d_stock_kernel<16>, smem) : set_smem(solve1d_stock_kernel<8>, smem);
}

// instantiated variants of the slice-streaming 2D kernel: field replicas G, deposit replicas D, threads of the one block per SM
typedef void (*p2d_kernel_t)(P2DArgs);
static p2d_kernel_t stream_kernel(int G, int D, int threads)
{
    if (G == 4 && D == 8 && threads == 512) return particles_2d3v_stream<4, 8, 512>;
    if (G == 1 && D == 1 && threads == 512) return particles_2d3v_stream<1, 1, 512>; // measurement: what the replicas buy
    return nullptr;
}

template <typename K>
static int occupancy_blocks(K kernel, int threads, size_t smem, int sms, int64_t work, int *out)
{
    int per_sm = 0;
    PG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) return fail(PICGOLF_ERR_CUDA, "kernel does not fit on an SM (smem %zu)", smem);
    int64_t want = (work + threads - 1) / threads;
    int64_t cap = (int64_t)per_sm * sms;
    *out = (int)std::max<int64_t>(1, std::min(want, cap));
    return 0;
}

// ------------------------------------------------------------------------------------------
// library
// ------------------------------------------------------------------------------------------
PG_API int picgolf_version(void) { return PICGOLF_VERSION; }
PG_API const char *picgolf_last_error(void) { return g_err; }
PG_API int picgolf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

PG_API int picgolf_config_default(picgolf_config *c, int scheme, int quiet)
{
    if (!c) return fail(PICGOLF_ERR_ARG, "cfg is NULL");
    memset(c, 0, sizeof(*c));
    c->struct_size = (int32_t)sizeof(*c);
    c->scheme = scheme;
    c->half_width = 6; c->max_sweeps = 10; c->diag_every = 1;
    c->device = -1; c->rank = 0; c->nranks = 1; c->local_first = -1; c->local_count = -1;
    const double pi = 3.14159265358979323846;
    switch (scheme) {
    case PICGOLF_NGP_LEAPFROG: // NGPFourier.jl:1
        c->N = 128; c->P = 64 * c->N; c->dt = 1.0 / (4 * c->N); c->T = 1024; c->W = 200;
        c->w = c->W / (double)c->P * (double)c->N;
        break;
    case PICGOLF_GAUSS_LEAPFROG: // Gaussian.jl:2   (deposit scale w/dx)
        c->N = 128; c->P = 64 * c->N; c->dt = 1.0 / (10 * c->N); c->T = 1024; c->W = 1600;
        c->w = c->W / (double)c->P / (1.0 / (double)c->N);
        break;
    case PICGOLF_AREA_SIMPSON13:  // AreaFixedPointQuietSimpson13.jl:1-4 (l=1e-14 is set after the switch)
    case PICGOLF_GAUSS_SIMPSON13: // GaussianFixedPointQuietSimpson13.jl:1-6 (same literals as the quiet fixed point)
        quiet = 1;
        /* fall through */
    case PICGOLF_GAUSS_FIXEDPOINT:
        if (!quiet) { // GaussianFixedPoint.jl:1-5
            c->N = 128; c->P = 32 * c->N; c->dt = 1.0 / (6 * c->N); c->T = 1024; c->W = 400;
            c->rtol = 1e-8; c->atol = 0; c->half_width = 6;
        } else { // GaussianFixedPointQuiet.jl:1-6
            c->N = 64; c->P = 32 * c->N; c->dt = 1.0 / (6 * c->N); c->T = 1 << 13; c->W = 32 * pi * pi / 3;
            c->rtol = 4 * 2.220446049250313e-16; c->atol = 0; c->half_width = 7;
        }
        c->w = c->W / (double)c->P * (double)c->N;
        break;
    case PICGOLF_GAUSS_BORIS_1D2V: { // NGP1D2V.jl:22-23
        c->N = 512; c->P = 15 * c->N; c->T = (1 << 14) / 16; c->diag_every = 16; c->half_width = 7;
        double n0 = 4 * pi * pi, vth = sqrt(n0) / (double)c->N / 4;
        c->W = n0; c->dt = 1 / (double)c->N / (6 * vth); c->B0 = sqrt(n0) / 16; c->w = n0 / (double)c->P;
        break;
    }
    case PICGOLF_GAUSS_BORIS_1D2V2S: { // NGP1D2V2S.jl:13-14
        c->N = 256; c->P = 8 * c->N; c->T = (1 << 16) / 32; c->diag_every = 32; c->half_width = 7; c->mass_ratio = 8; // T=2^16; TO=T/32 rows of T/TO=32 steps
        double n0 = 4 * pi * pi, vth = sqrt(n0) / (double)c->N / 8;
        c->W = n0; c->dt = 1 / (double)c->N / (16 * vth); c->B0 = sqrt(n0) / 8; c->w = n0 / (double)(2 * c->P);
        break;
    }
    case PICGOLF_CIC_BORIS_2D3V: { // Electrostatic2D3V.jl:23-25
        c->N = 128; c->NY = 128; c->P = c->N * c->NY * 32; c->T = 1 << 13;
        double NG = sqrt((double)(c->N * c->N + c->NY * c->NY));
        double n0 = 4 * pi * pi, vth = sqrt(n0) / NG;
        c->W = n0; c->dt = 1 / NG / (6 * vth); c->B0 = sqrt(n0) / 4; c->diag_every = 2;
        c->w = n0 / (double)c->P / ((1.0 / (double)c->N) * (1.0 / (double)c->NY));
        break;
    }
    default:
        return fail(PICGOLF_ERR_ARG, "unknown scheme %d", scheme);
    }
    if (scheme == PICGOLF_AREA_SIMPSON13) { c->rtol = 1e-14; c->half_width = 1; }
    return 0;
}

// ------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------
static void install_set(picgolf_handle h, int s);

static int destroy_impl(picgolf_handle h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->up_stream) { cudaStreamSynchronize(h->up_stream); cudaStreamSynchronize(h->down_stream); }
    if (h->sset_n) { // the ring of picgolf_step_streamed: set 0 is freed below with the handle's own arrays
        install_set(h, 0);
        for (int q = 1; q < 3; ++q) for (int i = 0; i < h->sset_n; ++i) if (h->sset[q].a[i]) cudaFree(h->sset[q].a[i]);
        for (int q = 0; q < 3; ++q) { if (h->ev_up[q]) cudaEventDestroy(h->ev_up[q]); if (h->ev_comp[q]) cudaEventDestroy(h->ev_comp[q]); if (h->ev_down[q]) cudaEventDestroy(h->ev_down[q]); }
    }
    if (h->up_stream) cudaStreamDestroy(h->up_stream);
    if (h->down_stream) cudaStreamDestroy(h->down_stream);
    h->timer.destroy();
    for (auto &g : h->step_graph) if (g) cudaGraphExecDestroy(g);
    for (auto &g : h->loop_graph) if (g.second) cudaGraphExecDestroy(g.second);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->slow_host) cudaFreeHost(h->slow_host);
    if (h->slow_ev) cudaEventDestroy(h->slow_ev);
    for (auto &e : h->run_ev) if (e) cudaEventDestroy(e);
    if (h->comm && nccl::CommDestroy) nccl::CommDestroy(h->comm);
    for (int q = 0; q < PEER_MAX; ++q) if (h->peer_ptr[q] && q != h->rank) cudaIpcCloseMemHandle(h->peer_ptr[q]);
    if (h->peer_mine) cudaFree(h->peer_mine);
    if (h->peer_err) cudaFree(h->peer_err);
    void *ptrs[] = {h->xb[0], h->xb[1], h->vb[0], h->vb[1], h->p2[0][0], h->p2[0][1], h->p2[0][2], h->p2[0][3], h->p2[0][4],
                    h->p2[1][0], h->p2[1][1], h->p2[1][2], h->p2[1][3], h->p2[1][4], h->bin_start, h->item_off, h->vy1, h->hist,
                    h->rho_last, h->E, h->rho_base[0] ? nullptr : (void *)h->rho_fx, h->rho_base[0], h->rho_base[1],
                    h->tw, h->twy, h->Z, h->E2, h->epartials, h->ctrl, h->partials, h->raw,
                    h->pid[0], h->pid[1], h->bin_count, h->bin_cursor, h->slow_count, h->Gpoly, h->Mg, h->snap[0], h->snap[1], h->snap[2],
                    h->dft_spec, h->dft_part, h->dft_arrive, h->dft_tw, h->vspare, h->fs_sync};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

PG_API int picgolf_destroy(picgolf_handle h) { return destroy_impl(h); }

static int create_impl(const picgolf_config *cfg, picgolf_handle h)
{
    const picgolf_config &c = h->cfg;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(PICGOLF_ERR_CUDA, "no CUDA device: libpicgolf has no CPU path");
    }
    if (c.device >= 0) h->device = c.device; else PG_CUDA(cudaGetDevice(&h->device));
    if (h->device >= ndev) return fail(PICGOLF_ERR_ARG, "device %d out of range (%d visible)", h->device, ndev);
    PG_CUDA(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    PG_CUDA(cudaGetDeviceProperties(&prop, h->device));
    h->sms = prop.multiProcessorCount;
    PG_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));

    { const char *e = getenv("PICGOLF_LOOP"); h->loop_off = e && e[0] == '0'; } // A/B: fixed schedule instead of the device-driven loop
    h->is2d = c.scheme == PICGOLF_CIC_BORIS_2D3V;
    h->simpson = c.scheme == PICGOLF_GAUSS_SIMPSON13 || c.scheme == PICGOLF_AREA_SIMPSON13;
    h->fixedpoint = c.scheme == PICGOLF_GAUSS_FIXEDPOINT || h->simpson;
    h->ngp = c.scheme == PICGOLF_NGP_LEAPFROG;
    h->b1d2v = c.scheme == PICGOLF_GAUSS_BORIS_1D2V || c.scheme == PICGOLF_GAUSS_BORIS_1D2V2S;
    const int64_t total = c.scheme == PICGOLF_GAUSS_BORIS_1D2V2S ? 2 * c.P : c.P; // particles held by all ranks together
    h->nranks = std::max(1, c.nranks);
    h->rank = c.rank;
    if (h->rank < 0 || h->rank >= h->nranks) return fail(PICGOLF_ERR_ARG, "rank %d outside [0,%d)", h->rank, h->nranks);
    if (c.local_first >= 0 && c.local_count >= 0) { h->first = c.local_first; h->count = c.local_count; }
    else { // even split, remainder to the low ranks
        int64_t q = total / h->nranks, r = total % h->nranks;
        h->first = q * h->rank + std::min<int64_t>(h->rank, r);
        h->count = q + (h->rank < r ? 1 : 0);
    }
    if (h->first < 0 || h->first + h->count > total) return fail(PICGOLF_ERR_ARG, "local shard outside [0,P)");
    h->T = std::max<int64_t>(1, c.T);

    const size_t n = (size_t)h->count;
    if (!h->is2d) {
        const int N = (int)c.N;
        h->ncell = N;
        PG_TRY(dalloc(&h->xb[0], n)); PG_TRY(dalloc(&h->vb[0], n));
        if (h->fixedpoint) { PG_TRY(dalloc(&h->xb[1], n)); PG_TRY(dalloc(&h->vb[1], n)); }
        const int rows = h->simpson ? 3 : 1; // Simpson-1/3: E1|E2|E3 and rho1|rho2|rho3
        h->grid_rows = rows;
        PG_TRY(dalloc(&h->rho_last, N)); PG_TRY(dalloc(&h->E, (size_t)rows * N));
        PG_TRY(dalloc(&h->rho_fx, (size_t)rows * N + 1)); // + 1: the flush counter rides along with the NCCL all-reduce
        if (h->fixedpoint && !h->simpson) {
            h->rho_base[0] = h->rho_fx;
            PG_TRY(dalloc(&h->rho_base[1], (size_t)N + 1));
            PG_CUDA(cudaMemset(h->rho_base[1], 0, ((size_t)N + 1) * sizeof(unsigned long long)));
            h->rho_next = h->rho_base[1];
        }
        PG_CUDA(cudaMemset(h->rho_last, 0, N * sizeof(double)));
        PG_CUDA(cudaMemset(h->E, 0, (size_t)rows * N * sizeof(double)));
        PG_CUDA(cudaMemset(h->rho_fx, 0, ((size_t)rows * N + 1) * sizeof(unsigned long long)));
        if (!h->ngp) {
            // fixed-point format of the Gaussian deposit grid: weights are <= 1 and sum to 1 per particle, so
            // no cell can exceed P (all ranks) -> 62 - ceil(log2(P+1)) fractional bits can never overflow.
            int frac = std::max(8, std::min(60, 62 - ilog2(total + 1)));
            h->fx_scale = ldexp(1.0, frac); h->fx_inv = ldexp(1.0, -frac);
        }
        PG_TRY(make_twiddles(&h->tw, N));
        h->smem_pass = (size_t)(2 * N + 32) * sizeof(double);
        h->npart = 2;
        PG_TRY(set_smem_solve1d((int)N));
        h->dft = !is_pow2(N); // NGP leapfrog on a grid that is not 2^k (validated in picgolf_create): direct transforms
        if (h->dft) {
            PG_TRY(dalloc(&h->dft_spec, (size_t)N)); PG_TRY(dalloc(&h->dft_part, (size_t)dft_blocks((int)N))); PG_TRY(dalloc(&h->dft_arrive, 1));
            PG_TRY(make_dft_twiddles(&h->dft_tw, (int)N));
            PG_CUDA(cudaMemset(h->dft_arrive, 0, sizeof(unsigned int)));
            PG_TRY(set_smem(solve1d_dft_fwd, dft_smem_bytes((int)N))); PG_TRY(set_smem(solve1d_dft_inv, dft_smem_bytes((int)N)));
        }
        if (h->b1d2v) {
            PG_TRY(dalloc(&h->vy1, n));
            PG_TRY(dalloc(&h->hist, (size_t)N * h->T));
            PG_CUDA(cudaMemset(h->hist, 0, (size_t)N * h->T * sizeof(double)));
            h->smem_b1 = (size_t)(2 * N + 32) * 8;
            h->npart = 3;
            PG_TRY(set_smem(b1d2v_pass, h->smem_b1));
            PG_TRY(occupancy_blocks(b1d2v_pass, PG_THREADS, h->smem_b1, h->sms, h->count, &h->nblocks));
        } else if (h->simpson) {
            h->smem_sp1 = (size_t)3 * N * 8;
            h->smem_spk = (size_t)(4 * N + 32) * 8;
            PG_TRY(set_smem(sp_pass0<0>, (size_t)N * 8)); PG_TRY(set_smem(sp_pass0<1>, (size_t)N * 8));
            PG_TRY(set_smem(sp_pass1<0>, h->smem_sp1)); PG_TRY(set_smem(sp_pass1<1>, h->smem_sp1));
            PG_TRY(set_smem(sp_passk<0>, h->smem_spk)); PG_TRY(set_smem(sp_passk<1>, h->smem_spk));
            PG_TRY(set_smem(solve_simpson23_kernel, h->smem_pass));
            PG_TRY(occupancy_blocks(sp_passk<0>, PG_THREADS, h->smem_spk, h->sms, h->count, &h->nblocks));
        } else if (h->fixedpoint) {
            PG_TRY(set_smem(fp_pass_atomic<true>, h->smem_pass));
            PG_TRY(set_smem(fp_pass_atomic<false>, h->smem_pass));
            PG_TRY(occupancy_blocks(fp_pass_atomic<false>, PG_THREADS, h->smem_pass, h->sms, h->count, &h->nblocks));
            const int64_t ppc = h->count / N;
            // deterministic: per-deposit fixed point in the order-free atomic kernel (bit-identical for any
            // particle order, run and GPU count); the sorted path is reproducible only between sweeps.
            if (c.deterministic) h->sorted = false;
            else if (c.deposit_mode == PICGOLF_DEPOSIT_SORTED) h->sorted = true;
            else if (c.deposit_mode == PICGOLF_DEPOSIT_AUTO) h->sorted = h->count >= (1 << 18) && ppc >= 64;
            // many particles per cell: sub-cell polynomial passes (HBM-bound instead of FP64-bound); with deterministic = 1
            // their lanes sum integers, so the fast path stays bit-reproducible (the windowed fp_pass_sorted does not)
            if (c.deposit_mode == PICGOLF_DEPOSIT_POLY || (c.deposit_mode == PICGOLF_DEPOSIT_AUTO && h->count >= (1 << 22) && ppc >= 1024)) {
                h->sorted = true; h->poly = true; h->det = c.deterministic != 0;
            }
            if (h->sorted) {
                h->K = (int)std::max<int64_t>(1, std::min<int64_t>(64, ppc / 16));
                h->K = (h->K + SORTED_NP - 1) / SORTED_NP * SORTED_NP; // whole groups of SORTED_NP batches
                // polynomial mode: (cell, sign v, sub-cell position) bins, at most POLY_MAXBINS and >= 512 particles each
                h->sublg = 0;
                if (h->poly) {
                    while (h->sublg < POLY_SUBLG && ((int64_t)2 * N << (h->sublg + 1)) <= POLY_MAXBINS &&
                           h->count / ((int64_t)2 * N << (h->sublg + 1)) >= 512) ++h->sublg;
                }
                h->nbins = h->poly ? (2 * N) << h->sublg : N;
                // re-sort before the slowest/fastest particles (|v| ~ 3) have drifted ~5 cells from their bin
                double cells_per_step = 3.0 * c.dt * (double)N;
                h->sort_every = c.sort_every > 0 ? c.sort_every : (int)std::max(1.0, std::min(1000.0, floor(5.0 / cells_per_step)));
                h->sort_auto = c.sort_every <= 0;
                h->smem_sorted = (size_t)(PG_THREADS / 32) * WIN_WARP_DOUBLES * sizeof(double);
                PG_TRY(set_smem(fp_pass_sorted<true, SORTED_NP>, h->smem_sorted));
                PG_TRY(set_smem(fp_pass_sorted<false, SORTED_NP>, h->smem_sorted));
                int64_t warps = (h->count + 32LL * h->K - 1) / (32LL * h->K);
                PG_TRY(occupancy_blocks(fp_pass_sorted<false, SORTED_NP>, PG_THREADS, h->smem_sorted, h->sms, warps * 32, &h->nblocks_sorted));
                h->nblocks = std::max(h->nblocks, h->nblocks_sorted);
                if (h->poly) {
                    if (c.sort_every <= 0) h->sort_every = 16; // starting point; adapted from the flush counter every step
                    h->smem_poly = cp_smem_bytes(CP_THREADS);
                    PG_TRY(set_smem(fp_pass_poly<true, false>, h->smem_poly)); PG_TRY(set_smem(fp_pass_poly<false, false>, h->smem_poly));
                    PG_TRY(set_smem(fp_pass_poly<true, true>, h->smem_poly)); PG_TRY(set_smem(fp_pass_poly<false, true>, h->smem_poly));
                    PG_TRY(set_smem(mom2rho_kernel, CPM_SMEM));
                    PG_TRY(dalloc(&h->Gpoly, (size_t)CP_GS * CP_NSUB * N)); PG_TRY(dalloc(&h->Mg, (size_t)2 * CP_NC * CP_NSUB * N)); // x2: deterministic mode
                    PG_CUDA(cudaMemset(h->Gpoly, 0, (size_t)CP_GS * CP_NSUB * N * sizeof(double)));
                    PG_CUDA(cudaMemset(h->Mg, 0, (size_t)2 * CP_NC * CP_NSUB * N * sizeof(unsigned long long)));
                    // one contiguous range of >= 16 rows (of 64 particles) per warp
                    const int64_t rows = (h->count + 63) / 64;
                    PG_TRY(occupancy_blocks(fp_pass_poly<false, true>, CP_THREADS, h->smem_poly, h->sms, (rows + 15) / 16 * 32, &h->nblocks_poly));
                    h->nblocks = std::max(h->nblocks, h->nblocks_poly);
                    const char *fsenv = getenv("PICGOLF_FUSED_SORT"); // 0: every re-sort is the stand-alone counting sort (measurement / fallback)
                    h->fs_enabled = !(fsenv && atoi(fsenv) == 0) && (((int64_t)N << h->sublg) < ((int64_t)1 << 30));
                    if ((h->nbins + FS_CHUNK - 1) / FS_CHUNK > FS_MAXBLOCKS || h->nbins % 4) h->fs_enabled = false;
                    if (h->fs_enabled) {
                        PG_TRY(dalloc(&h->vspare, n));
                        PG_TRY(dalloc(&h->fs_sync, FS_MAXBLOCKS));
                        PG_CUDA(cudaMemset(h->fs_sync, 0, FS_MAXBLOCKS * sizeof(unsigned long long)));
                    }
                }
                PG_TRY(dalloc(&h->pid[0], n)); PG_TRY(dalloc(&h->pid[1], n));
                PG_TRY(dalloc(&h->bin_count, h->nbins)); PG_TRY(dalloc(&h->bin_cursor, h->nbins));
                PG_CUDA(cudaMemset(h->bin_count, 0, h->nbins * sizeof(unsigned int)));
                PG_TRY(dalloc(&h->slow_count, 1));
                PG_CUDA(cudaMemset(h->slow_count, 0, sizeof(unsigned long long)));
                if (!h->poly) {
                    PG_TRY(set_smem(sort_hist_kernel, (size_t)h->nbins * 4));
                    PG_TRY(set_smem(sort_scatter_kernel<2>, (size_t)h->nbins * 8));
                }
            }
        } else if (h->ngp) {
            h->smem_lf = lf_smem_bytes(0, N);
            PG_TRY(set_smem(lf_pass<0, false>, h->smem_lf)); PG_TRY(set_smem(lf_pass<0, true>, h->smem_lf));
            PG_TRY(occupancy_blocks(lf_pass<0, true>, PG_THREADS, h->smem_lf, h->sms, (h->count + 3) / 4, &h->nblocks));
            // large shards stream their particle tiles with TMA bulk copies (one persistent block per SM)
            h->ngp_tma = c.deposit_mode != PICGOLF_DEPOSIT_ATOMIC && h->count >= (1 << 20) && lf_tma_smem_bytes(N) <= 200 * 1024;
            if (h->ngp_tma) {
                h->smem_lf = lf_tma_smem_bytes(N);
                PG_TRY(set_smem(lf_pass_ngp_tma<false>, h->smem_lf)); PG_TRY(set_smem(lf_pass_ngp_tma<true>, h->smem_lf));
                h->nblocks = h->sms;
            }
        } else {
            h->smem_lf = lf_smem_bytes(1, N);
            PG_TRY(set_smem(lf_pass<1, false>, h->smem_lf)); PG_TRY(set_smem(lf_pass<1, true>, h->smem_lf));
            PG_TRY(occupancy_blocks(lf_pass<1, true>, PG_THREADS, h->smem_lf, h->sms, h->count, &h->nblocks));
        }
    } else {
        const int NX = (int)c.N, NY = (int)c.NY;
        h->ncell = (int64_t)NX * NY;
        for (int q = 0; q < 5; ++q) PG_TRY(dalloc(&h->p2[0][q], n));
        PG_TRY(dalloc(&h->rho_fx, h->ncell)); PG_TRY(dalloc(&h->rho_last, h->ncell));
        PG_TRY(dalloc(&h->Z, h->ncell)); PG_TRY(dalloc(&h->E2, h->ncell));
        PG_TRY(dalloc(&h->epartials, NY / ROWS_PER_BLOCK));
        PG_CUDA(cudaMemset(h->rho_fx, 0, h->ncell * sizeof(unsigned long long)));
        PG_CUDA(cudaMemset(h->rho_last, 0, h->ncell * sizeof(double)));
        PG_CUDA(cudaMemset(h->E2, 0, h->ncell * sizeof(double2)));
        if (c.field_history)
            for (int w = 0; w < 3; ++w) {
                PG_TRY(dalloc(&h->snap[w], (size_t)h->ncell * h->T));
                PG_CUDA(cudaMemset(h->snap[w], 0, (size_t)h->ncell * h->T * sizeof(double)));
            }
        PG_TRY(make_twiddles(&h->tw, NX)); PG_TRY(make_twiddles(&h->twy, NY));
        PG_TRY(set_smem(solve2d_rows_fwd, (size_t)2 * ROWS_PER_BLOCK * NX * 8));
        PG_TRY(set_smem(solve2d_rows_inv, (size_t)(2 * ROWS_PER_BLOCK * NX + 32) * 8));
        PG_TRY(set_smem(solve2d_cols, (size_t)2 * COLS_PER_BLOCK * (NY + 1) * 8));
        h->npart = 3;
        PG_TRY(occupancy_blocks(particles_2d3v_kernel, PG_THREADS, 0, h->sms, h->count, &h->nblocks));
        int ilog2f = 0;
        {   // fixed-point formats: CIC weights are <= 1 and sum to 1 per particle (no overflow possible)
            int frac = std::max(8, std::min(60, 62 - ilog2(c.P + 1)));
            ilog2f = frac;
            h->fx_scale = ldexp(1.0, frac); h->fx_inv = ldexp(1.0, -frac);
            int fracw = std::max(frac, 62 - ilog2((int64_t)T2_CHUNK + 1)); // a window only sees T2_CHUNK particles
            h->fxw_scale = ldexp(1.0, fracw); h->fx_shift = fracw - frac;
        }
        const int64_t ppc = h->count / h->ncell;
        if (c.deterministic) h->sorted = false;
        else if (c.deposit_mode == PICGOLF_DEPOSIT_SORTED) h->sorted = true;
        else if (c.deposit_mode == PICGOLF_DEPOSIT_AUTO) h->sorted = h->count >= (1 << 20) && ppc >= 8;
        if (h->sorted) {
            const int ntx = std::max(1, NX >> T2_SHIFT), nty = std::max(1, NY >> T2_SHIFT);
            h->nbins = ntx * nty;
            h->sort_every = c.sort_every > 0 ? c.sort_every : 16; // starting point; adapted from the slow-path counter
            h->sort_auto = c.sort_every <= 0;
            for (int q = 0; q < 5; ++q) PG_TRY(dalloc(&h->p2[1][q], n));
            PG_TRY(dalloc(&h->pid[0], n)); PG_TRY(dalloc(&h->pid[1], n));
            PG_TRY(dalloc(&h->bin_count, h->nbins)); PG_TRY(dalloc(&h->bin_cursor, h->nbins));
            PG_TRY(dalloc(&h->bin_start, h->nbins)); PG_TRY(dalloc(&h->item_off, h->nbins + 1));
            PG_CUDA(cudaMemset(h->bin_count, 0, h->nbins * sizeof(unsigned int)));
            PG_TRY(dalloc(&h->slow_count, 1));
            PG_CUDA(cudaMemset(h->slow_count, 0, sizeof(unsigned long long)));
            PG_TRY(set_smem(sort_hist_kernel, (size_t)h->nbins * 4));
            PG_TRY(set_smem(sort_scatter_kernel<5>, (size_t)h->nbins * 8));
            int64_t items = h->count / T2_CHUNK + h->nbins;
            PG_TRY(occupancy_blocks(particles_2d3v_tiled, PG_THREADS, 0, h->sms, items * PG_THREADS, &h->nblocks_sorted));
            // particle kernel of the tile-sorted path (PICGOLF_2D_KERNEL): "stream" (default: slice streaming with replicated
            // windows), "stream11_512" (the same without replicas) or "tiled" (8192-particle work items, plain loads)
            const char *kv = getenv("PICGOLF_2D_KERNEL");
            h->k2d = 2; h->k2d_a = 4; h->k2d_b = 8; h->k2d_c = 512;
            if (kv && !strcmp(kv, "tiled")) h->k2d = 0;
            else if (kv && !strncmp(kv, "stream", 6) && strlen(kv) > 9) { h->k2d_a = kv[6] - '0'; h->k2d_b = kv[7] - '0'; h->k2d_c = atoi(kv + 9); }
            else if (kv && strcmp(kv, "stream")) return fail(PICGOLF_ERR_ARG, "PICGOLF_2D_KERNEL=%s: no such kernel", kv);
            if (h->k2d == 2) {
                p2d_kernel_t kern = stream_kernel(h->k2d_a, h->k2d_b, h->k2d_c);
                if (!kern) return fail(PICGOLF_ERR_ARG, "PICGOLF_2D_KERNEL=%s: no such variant", kv);
                h->smem_ring = s2_smem_bytes(h->k2d_a, h->k2d_b, h->k2d_c);
                PG_TRY(set_smem(kern, h->smem_ring));
                h->nblocks_sorted = h->sms; // one block per SM, each streams a contiguous slice of the sorted arrays
                // its deposit window is flushed every S2_FLUSH particles instead of every T2_CHUNK
                int fracs = std::max(ilog2f, 62 - ilog2((int64_t)S2_FLUSH + 1));
                h->fxw_scale = ldexp(1.0, fracs); h->fx_shift = fracs - ilog2f;
            }
            h->nblocks = std::max(h->nblocks, h->nblocks_sorted);
        }
    }
    PG_TRY(dalloc(&h->ctrl, 1));
    Ctrl c0; memset(&c0, 0, sizeof(c0)); c0.final_k = -1;
    PG_CUDA(cudaMemcpy(h->ctrl, &c0, sizeof(c0), cudaMemcpyHostToDevice));
    PG_TRY(dalloc(&h->partials, (size_t)4 * h->nblocks)); // up to 3 doubles per block; 4 integers in deterministic polynomial mode
    PG_CUDA(cudaMemset(h->partials, 0, (size_t)4 * h->nblocks * sizeof(double)));
    PG_TRY(dalloc(&h->raw, (size_t)4 * h->T));
    PG_CUDA(cudaMemset(h->raw, 0, (size_t)4 * h->T * sizeof(double)));
    (void)cfg;
    return 0;
}

PG_API int picgolf_create(const picgolf_config *cfg, picgolf_handle *out)
{
    if (!cfg || !out) return fail(PICGOLF_ERR_ARG, "NULL argument");
    if (cfg->struct_size != (int32_t)sizeof(picgolf_config))
        return fail(PICGOLF_ERR_ARG, "struct_size %d != %zu (header/library mismatch)", cfg->struct_size, sizeof(picgolf_config));
    const picgolf_config &c = *cfg;
    if (c.scheme < PICGOLF_NGP_LEAPFROG || c.scheme > PICGOLF_GAUSS_BORIS_1D2V2S) return fail(PICGOLF_ERR_ARG, "unknown scheme %d", c.scheme);
    if (c.scheme == PICGOLF_GAUSS_BORIS_1D2V2S && !(c.mass_ratio > 0)) return fail(PICGOLF_ERR_ARG, "mass_ratio must be positive");
    if (c.P < 1) return fail(PICGOLF_ERR_ARG, "P must be >= 1");
    if (!(c.dt > 0) || !isfinite(c.dt)) return fail(PICGOLF_ERR_ARG, "dt must be positive and finite");
    if (!isfinite(c.w)) return fail(PICGOLF_ERR_ARG, "w must be finite");
    if (c.scheme == PICGOLF_CIC_BORIS_2D3V) {
        if (c.N < 16 || c.NY < 16 || c.N > 1024 || c.NY > 1024) return fail(PICGOLF_ERR_ARG, "2D grid must be 16..1024 per side");
        if (!is_pow2(c.N) || !is_pow2(c.NY))
            return fail(PICGOLF_ERR_UNSUPPORTED, "NX=%lld NY=%lld: only power-of-two grids are built (radix-2 shared-memory FFT)", (long long)c.N, (long long)c.NY);
    } else {
        if (c.N < 16 || c.N > 8192) return fail(PICGOLF_ERR_ARG, "N must be in 16..8192");
        // grids that are not 2^k: the NGP leapfrog only (NGPFourier.jl's fft takes any N; its ik vector needs an even one) -- direct
        // transforms (solve1d_dft_*).  The erf-shape kernels rest on delta = c*N - round(c*N) being exact, which needs N = 2^k.
        if (!is_pow2(c.N) && (c.scheme != PICGOLF_NGP_LEAPFROG || (c.N & 1)))
            return fail(PICGOLF_ERR_UNSUPPORTED, "N=%lld: grids that are not a power of two are built for the NGP leapfrog and even N only", (long long)c.N);
        if (c.scheme != PICGOLF_NGP_LEAPFROG && c.scheme != PICGOLF_AREA_SIMPSON13 && c.half_width != 6 && c.half_width != 7)
            return fail(PICGOLF_ERR_ARG, "half_width must be 6 or 7");
        if ((c.scheme == PICGOLF_GAUSS_SIMPSON13 || c.scheme == PICGOLF_AREA_SIMPSON13) && c.N > 4096) return fail(PICGOLF_ERR_ARG, "Simpson-1/3 scheme: N must be <= 4096");
        if ((c.scheme == PICGOLF_GAUSS_FIXEDPOINT || c.scheme == PICGOLF_GAUSS_SIMPSON13 || c.scheme == PICGOLF_AREA_SIMPSON13) && (c.max_sweeps < 1 || c.max_sweeps > 64))
            return fail(PICGOLF_ERR_ARG, "max_sweeps must be in 1..64");
    }
    picgolf_handle h = new picgolf_handle_s();
    h->cfg = c;
    if (h->cfg.diag_every < 1) h->cfg.diag_every = 1;
    int rc = create_impl(cfg, h);
    if (rc != 0) { char keep[512]; memcpy(keep, g_err, sizeof(keep)); destroy_impl(h); memcpy(g_err, keep, sizeof(keep)); return rc; }
    *out = h;
    return 0;
}

PG_API int picgolf_local_range(picgolf_handle h, int64_t *first, int64_t *count)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (first) *first = h->first;
    if (count) *count = h->count;
    return 0;
}

// ------------------------------------------------------------------------------------------
// particle state
// ------------------------------------------------------------------------------------------
// Launches of the device-driven loop are counted on the device (one counter increment per sweep); fold them into the
// host-side count before Ctrl is cleared.
static int bank_loop_launches(picgolf_handle h)
{
    if (h->loop_steps == 0) return 0;
    PG_CUDA(cudaStreamSynchronize(h->stream));
    Ctrl c;
    PG_CUDA(cudaMemcpy(&c, h->ctrl, sizeof(c), cudaMemcpyDeviceToHost));
    h->launches += (int64_t)c.loop_sweeps * h->loop_launches_per_sweep + (int64_t)c.loop_fs_sweeps;
    h->loop_steps = 0;
    return 0;
}

static int reset_run_state(picgolf_handle h)
{
    PG_TRY(bank_loop_launches(h));
    if (h->sset_n) { // back from picgolf_step_streamed to the plain calls: drain the copy streams (the caller has just enqueued
                     // its upload into the installed set on h->stream, which is ordered after that set's last step)
        PG_CUDA(cudaStreamSynchronize(h->up_stream)); PG_CUDA(cudaStreamSynchronize(h->down_stream));
        for (auto &u : h->sset_used) u = false;
    }
    Ctrl c0; memset(&c0, 0, sizeof(c0)); c0.final_k = -1;
    PG_CUDA(cudaMemcpyAsync(h->ctrl, &c0, sizeof(c0), cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(cudaMemsetAsync(h->rho_last, 0, h->ncell * sizeof(double), h->stream));
    if (h->rho_base[0]) { h->rho_fx = h->rho_base[0]; h->rho_next = h->rho_base[1]; }
    PG_CUDA(cudaMemsetAsync(h->rho_fx, 0, (size_t)h->grid_rows * h->ncell * sizeof(unsigned long long), h->stream));
    if (h->rho_next) PG_CUDA(cudaMemsetAsync(h->rho_next, 0, ((size_t)h->ncell + 1) * sizeof(unsigned long long), h->stream));
    if (h->is2d) PG_CUDA(cudaMemsetAsync(h->E2, 0, h->ncell * sizeof(double2), h->stream));
    else PG_CUDA(cudaMemsetAsync(h->E, 0, (size_t)h->grid_rows * h->ncell * sizeof(double), h->stream));
    if (h->Mg) PG_CUDA(cudaMemsetAsync(h->Mg, 0, (size_t)2 * CP_NC * CP_NSUB * h->ncell * sizeof(unsigned long long), h->stream));
    if (h->fs_sync) { // the epochs of the bin scan start over with the step counter; the bin counts are zero between steps
        PG_CUDA(cudaMemsetAsync(h->fs_sync, 0, FS_MAXBLOCKS * sizeof(unsigned long long), h->stream));
        PG_CUDA(cudaMemsetAsync(h->bin_count, 0, (size_t)h->nbins * sizeof(unsigned int), h->stream));
    }
    PG_CUDA(cudaStreamSynchronize(h->stream));
    h->par = 0; h->steps = 0; h->have_particles = true;
    h->pid_valid = false; h->pidpar = 0; h->since_sort = 0; h->have_deposit = false;
    h->slow_pending = false; h->steps_at_probe = 0; h->steps_at_probe_prev_steps = 0; h->force_sort = false; h->poly_quiet = false; h->probe_have_prev = false;
    for (auto &ps : h->probe_step) ps = -1;
    if (h->hist) PG_CUDA(cudaMemset(h->hist, 0, (size_t)h->ncell * h->T * sizeof(double)));
    for (int w = 0; w < 3; ++w)
        if (h->snap[w]) PG_CUDA(cudaMemset(h->snap[w], 0, (size_t)h->ncell * h->T * sizeof(double)));
    return 0;
}

PG_API int picgolf_set_particles(picgolf_handle h, const double *x, const double *v, int64_t count)
{
    if (!h || !x || !v) return fail(PICGOLF_ERR_ARG, "NULL argument");
    if (h->is2d) return fail(PICGOLF_ERR_ARG, "use picgolf_set_particles_2d3v for the 2D3V scheme");
    if (h->b1d2v) return fail(PICGOLF_ERR_ARG, "use picgolf_set_particles_1d2v for the 1D2V scheme");
    if (count != h->count) return fail(PICGOLF_ERR_ARG, "count %lld != local shard %lld", (long long)count, (long long)h->count);
    PG_TRY(use_device(h));
    PG_CUDA(cudaMemcpyAsync(h->xb[0], x, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(cudaMemcpyAsync(h->vb[0], v, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return reset_run_state(h);
}

PG_API int picgolf_set_particles_2d3v(picgolf_handle h, const double *x, const double *y, const double *vx,
                                      const double *vy, const double *vz, int64_t count)
{
    if (!h || !x || !y || !vx || !vy || !vz) return fail(PICGOLF_ERR_ARG, "NULL argument");
    if (!h->is2d) return fail(PICGOLF_ERR_ARG, "handle is not a 2D3V scheme");
    if (count != h->count) return fail(PICGOLF_ERR_ARG, "count %lld != local shard %lld", (long long)count, (long long)h->count);
    PG_TRY(use_device(h));
    const size_t b = count * sizeof(double);
    const double *src[5] = {x, y, vx, vy, vz};
    for (int q = 0; q < 5; ++q) PG_CUDA(cudaMemcpyAsync(h->p2[0][q], src[q], b, cudaMemcpyHostToDevice, h->stream));
    return reset_run_state(h);
}

PG_API int picgolf_set_particles_1d2v(picgolf_handle h, const double *x, const double *vx, const double *vy, int64_t count)
{
    if (!h || !x || !vx || !vy) return fail(PICGOLF_ERR_ARG, "NULL argument");
    if (!h->b1d2v) return fail(PICGOLF_ERR_ARG, "handle is not a 1D2V scheme");
    if (count != h->count) return fail(PICGOLF_ERR_ARG, "count %lld != local shard %lld", (long long)count, (long long)h->count);
    PG_TRY(use_device(h));
    const size_t b = count * sizeof(double);
    PG_CUDA(cudaMemcpyAsync(h->xb[0], x, b, cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(cudaMemcpyAsync(h->vb[0], vx, b, cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(cudaMemcpyAsync(h->vy1, vy, b, cudaMemcpyHostToDevice, h->stream));
    return reset_run_state(h);
}

PG_API int picgolf_get_particles_1d2v(picgolf_handle h, double *x, double *vx, double *vy, int64_t count)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (!h->b1d2v) return fail(PICGOLF_ERR_ARG, "handle is not a 1D2V scheme");
    if (!h->have_particles) return fail(PICGOLF_ERR_STATE, "particles were never set");
    if (count != h->count) return fail(PICGOLF_ERR_ARG, "count %lld != local shard %lld", (long long)count, (long long)h->count);
    PG_TRY(use_device(h));
    const size_t b = count * sizeof(double);
    if (x) PG_CUDA(cudaMemcpyAsync(x, h->xb[0], b, cudaMemcpyDeviceToHost, h->stream));
    if (vx) PG_CUDA(cudaMemcpyAsync(vx, h->vb[0], b, cudaMemcpyDeviceToHost, h->stream));
    if (vy) PG_CUDA(cudaMemcpyAsync(vy, h->vy1, b, cudaMemcpyDeviceToHost, h->stream));
    PG_CUDA(cudaStreamSynchronize(h->stream));
    return check_peer(h);
}

static int init_grid(picgolf_handle h) { return (int)std::max<int64_t>(1, std::min<int64_t>((h->count + 255) / 256, (int64_t)h->sms * 8)); }

PG_API int picgolf_init_quiet(picgolf_handle h)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (h->is2d || h->b1d2v) return fail(PICGOLF_ERR_ARG, "quiet start is a 1D1V initialisation (GaussianFixedPointQuiet.jl:2-3)");
    PG_TRY(use_device(h));
    quiet_start_kernel<<<init_grid(h), 256, 0, h->stream>>>(h->xb[0], h->vb[0], h->count, h->first, h->cfg.P);
    h->launches++;
    PG_CUDA(cudaGetLastError());
    return reset_run_state(h);
}

PG_API int picgolf_init_synthetic(picgolf_handle h, uint64_t seed, double vth)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (h->b1d2v) return fail(PICGOLF_ERR_UNSUPPORTED, "1D2V: pass x, vx, vy in with picgolf_set_particles_1d2v");
    PG_TRY(use_device(h));
    if (h->is2d)
        synthetic_2d3v_kernel<<<init_grid(h), 256, 0, h->stream>>>(h->p2[0][0], h->p2[0][1], h->p2[0][2], h->p2[0][3], h->p2[0][4], h->count,
                                                                   h->first, seed, vth);
    else
        synthetic_1d_kernel<<<init_grid(h), 256, 0, h->stream>>>(h->xb[0], h->vb[0], h->count, h->first, h->cfg.P, seed, vth);
    h->launches++;
    PG_CUDA(cudaGetLastError());
    return reset_run_state(h);
}

PG_API int picgolf_synchronize(picgolf_handle h)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    PG_TRY(use_device(h));
    if (h->up_stream) { PG_CUDA(cudaStreamSynchronize(h->up_stream)); }
    PG_CUDA(cudaStreamSynchronize(h->stream));
    if (h->down_stream) { PG_CUDA(cudaStreamSynchronize(h->down_stream)); }
    return check_peer(h);
}

PG_API int picgolf_get_particles(picgolf_handle h, double *x, double *v, int64_t count)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (h->is2d) return fail(PICGOLF_ERR_ARG, "use picgolf_get_particles_2d3v for the 2D3V scheme");
    if (!h->have_particles) return fail(PICGOLF_ERR_STATE, "particles were never set");
    if (count != h->count) return fail(PICGOLF_ERR_ARG, "count %lld != local shard %lld", (long long)count, (long long)h->count);
    PG_TRY(use_device(h));
    if (h->sorted && h->pid_valid) {
        // back to the caller's order through the free ping-pong buffer
        double *tmp = h->xb[1 - h->par];
        const double *src[2] = {h->xb[h->par], h->vb[h->par]};
        double *dst[2] = {x, v};
        for (int q = 0; q < 2; ++q) {
            if (!dst[q]) continue;
            unsort_kernel<<<h->sms * 8, 256, 0, h->stream>>>(src[q], h->pid[h->pidpar], tmp, h->count);
            h->launches++;
            PG_CUDA(cudaMemcpyAsync(dst[q], tmp, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        }
        PG_CUDA(cudaStreamSynchronize(h->stream));
        return check_peer(h);
    }
    if (x) PG_CUDA(cudaMemcpyAsync(x, h->xb[h->par], count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (v) PG_CUDA(cudaMemcpyAsync(v, h->vb[h->par], count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PG_CUDA(cudaStreamSynchronize(h->stream));
    return check_peer(h);
}

PG_API int picgolf_get_particles_2d3v(picgolf_handle h, double *x, double *y, double *vx, double *vy, double *vz,
                                      int64_t count)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (!h->is2d) return fail(PICGOLF_ERR_ARG, "handle is not a 2D3V scheme");
    if (!h->have_particles) return fail(PICGOLF_ERR_STATE, "particles were never set");
    if (count != h->count) return fail(PICGOLF_ERR_ARG, "count %lld != local shard %lld", (long long)count, (long long)h->count);
    PG_TRY(use_device(h));
    const size_t b = count * sizeof(double);
    double *dst[5] = {x, y, vx, vy, vz};
    for (int i = 0; i < 5; ++i) {
        if (!dst[i]) continue;
        const double *src = h->p2[h->par][i];
        if (h->sorted && h->pid_valid) { // back to the caller's order through a free ping-pong buffer
            double *tmp = h->p2[1 - h->par][0];
            unsort_kernel<<<h->sms * 8, 256, 0, h->stream>>>(src, h->pid[h->pidpar], tmp, h->count);
            h->launches++;
            src = tmp;
        }
        PG_CUDA(cudaMemcpyAsync(dst[i], src, b, cudaMemcpyDeviceToHost, h->stream));
    }
    PG_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// the loop body
// ------------------------------------------------------------------------------------------
static PeerArgs peer_args(picgolf_handle h)
{
    PeerArgs p;
    memset(&p, 0, sizeof(p));
    for (int q = 0; q < h->nranks; ++q) p.peer[q] = h->peer_ptr[q];
    p.nranks = h->nranks; p.rank = h->rank; p.ncell = h->ncell; p.error = h->peer_err;
    p.flush_src = h->slow_count;
    p.timeout_cycles = PEER_TIMEOUT_CYCLES;
    if (const char *ts = getenv("PICGOLF_PEER_TIMEOUT_S")) // test runs: give up quickly instead of holding a GPU box for minutes
        if (atof(ts) > 0) p.timeout_cycles = (long long)(atof(ts) * 1.9e9);
    return p;
}

static int allreduce_grid(picgolf_handle h, int row0 = 0, int nrows = 1)
{
    h->peer_this_solve = false;
    if (!h->comm) return 0;
    if (h->peer_ok && !h->is2d && !h->simpson && !h->dft && row0 == 0 && nrows == 1) {
        // 1D schemes: publish this rank's grid; the solve kernel that follows adds the ranks' grids up itself
        const int sp = h->timer.begin(ST_REDUCE, h->stream);
        peer_publish_kernel<<<1, 1024, 0, h->stream>>>(peer_args(h), h->rho_fx, &h->ctrl->final_k, h->fixedpoint ? 1 : 0);
        h->timer.end(sp, h->stream);
        h->launches++;
        h->peer_this_solve = true;
        return 0;
    }
    const int sp1_ = h->timer.begin(ST_REDUCE, h->stream);
    int rc;
    // integer grid: two's-complement sums are exact and order independent
    unsigned long long *g = h->rho_fx + (size_t)row0 * h->ncell;
    // polynomial mode: slot ncell carries this rank's flush counter (written by mom2rho_kernel) -> global sum for the probe
    const size_t extra = (h->poly && h->use_sorted_now && row0 == 0 && nrows == 1) ? 1 : 0;
    rc = nccl::AllReduce(g, g, (size_t)nrows * h->ncell + extra, nccl::Int64, nccl::Sum, h->comm, h->stream);
    h->timer.end(sp1_, h->stream);
    return nccl::check(rc, "ncclAllReduce(rho)");
}

static int launch_solve1d(picgolf_handle h, int k, bool simpson_e1 = false, cudaGraphConditionalHandle cond = 0)
{
    const picgolf_config &c = h->cfg;
    if (h->dft) {
        SolveDftArgs d;
        memset(&d, 0, sizeof(d));
        d.rho_fx = h->rho_fx; d.rho_last = h->rho_last; d.E = h->E; d.spec = h->dft_spec; d.part = h->dft_part; d.arrive = h->dft_arrive;
        d.ctrl = h->ctrl; d.w = c.w; d.fx_inv = h->fx_inv; d.N = (int)c.N; d.tw = h->dft_tw;
        const int nb = dft_blocks((int)c.N);
        const size_t sm = dft_smem_bytes((int)c.N);
        const int sp = h->timer.begin(ST_SOLVE, h->stream);
        solve1d_dft_fwd<<<nb, DFT_THREADS, sm, h->stream>>>(d);
        solve1d_dft_inv<<<nb, DFT_THREADS, sm, h->stream>>>(d);
        h->timer.end(sp, h->stream);
        h->launches += 2;
        return 0;
    }
    Solve1DArgs a;
    memset(&a, 0, sizeof(a));
    a.rho_in = nullptr; a.rho_fx = h->rho_fx; a.rho_last = h->rho_last; a.E = h->E; a.tw = h->tw; a.ctrl = h->ctrl;
    a.w = c.w; a.fx_inv = h->fx_inv; a.rtol = c.rtol; a.atol = c.atol;
    a.N = (int)c.N; a.lg = ilog2(c.N); a.fixedpoint = (h->fixedpoint && !simpson_e1) ? 1 : 0;
    a.k = k; a.max_sweeps = c.max_sweeps; a.store_normE1 = simpson_e1 ? 1 : 0;
    a.hist = nullptr; a.cond = cond;
    if (h->peer_this_solve) a.peer = peer_args(h);
    a.flush_slot = (h->comm && !h->peer_ok && h->poly && h->use_sorted_now) ? 1 : 0;
    if (h->b1d2v) { // Es[:,ti] .+= E with ti = cld(t, T/TO)   NGP1D2V.jl:56-57
        int64_t ti = h->steps / c.diag_every;
        if (ti < h->T) a.hist = h->hist + (size_t)ti * c.N;
    }
    const int sp2_ = h->timer.begin(ST_SOLVE, h->stream);
    launch_solve1d_kernel(a, h->stream);
    h->timer.end(sp2_, h->stream);
    h->launches++;
    return 0;
}

static bool poly_now(picgolf_handle h);
static int launch_step_end(picgolf_handle h, bool record)
{
    StepEndArgs a;
    a.partials = h->partials; a.epartials = h->is2d ? h->epartials : nullptr; a.raw = h->raw; a.ctrl = h->ctrl;
    a.nblocks = h->pass_blocks > 0 ? h->pass_blocks : h->nblocks; a.npart = h->npart; a.neblocks = h->is2d ? (int)(h->cfg.NY / ROWS_PER_BLOCK) : 0;
    a.T = (int)h->T; a.record = record ? 1 : 0; a.is2d = h->npart == 3 ? 1 : 0;
    a.det = (h->det && h->fixedpoint && !h->simpson && poly_now(h)) ? 1 : 0;
    step_end_kernel<<<1, 256, 0, h->stream>>>(a);
    h->launches++;
    return 0;
}

// Adaptive re-sort interval.  Called at every sort: if the previous probe of the slow-path counter has landed,
// compare the fraction of deposits that left their window since then with two thresholds and shorten / lengthen the
// interval; then start a new probe.  Never synchronises.
static void adapt_sort_interval(picgolf_handle h)
{
    if (!h->sort_auto || !h->slow_count) return;
    if (h->comm) return; // several GPUs: the probe below is timing dependent -- every rank keeps the same fixed interval,
                         // otherwise the ranks would sort at different steps and wait for each other's sorts
    if (!h->slow_host) {
        if (cudaMallocHost((void **)&h->slow_host, sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); h->sort_auto = false; return; }
        cudaEventCreateWithFlags(&h->slow_ev, cudaEventDisableTiming);
        *h->slow_host = 0;
    }
    if (h->slow_pending && cudaEventQuery(h->slow_ev) == cudaSuccess) {
        const unsigned long long now = *h->slow_host;
        const int64_t steps = std::max<int64_t>(1, h->steps_at_probe_prev_steps);
        const double frac = (double)(now - h->slow_seen) / ((double)h->count * (double)steps);
        h->slow_seen = now;
        if (frac > 1e-3) h->sort_every = std::max(2, h->sort_every / 2);
        else if (frac < 5e-5) h->sort_every = std::min(64, h->sort_every + std::max(1, h->sort_every / 2));
        h->slow_pending = false;
    } else {
        cudaGetLastError();
    }
    if (!h->slow_pending) {
        cudaMemcpyAsync(h->slow_host, h->slow_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
        cudaEventRecord(h->slow_ev, h->stream);
        h->slow_pending = true;
        h->steps_at_probe_prev_steps = h->steps - h->steps_at_probe;
        h->steps_at_probe = h->steps;
    }
}

// Polynomial mode: the flush counter is probed EVERY step: an 8-byte asynchronous copy into a ring of pinned slots at
// the start of the step.  The host runs at most POLY_RUNAHEAD steps ahead of the device (it waits for the end-of-step
// marker of an OLDER step, so the device queue never drains and no step is ever synchronised internally), and at step
// s it reads the slot filled at the start of step s - POLY_RUNAHEAD, which has certainly landed.  The decision is
// therefore a deterministic function of the step number and of the counter -- on several GPUs the counter is the SUM
// over the ranks (formed by the solve kernel next to the grid sum, pg_peer.cuh), so all ranks re-sort at the same steps
// and nobody waits for somebody else's sort.  Several times the count expected in sorted order means lanes alternate
// between cells all the time (warm beams shear a bin over several cells) -- then the next step re-sorts at once,
// whatever the interval, and the interval is shortened to what was survived.  Quiet intervals grow by half up to 64.
constexpr int POLY_RUNAHEAD = 4;
static void probe_poly_flushes(picgolf_handle h)
{
    if (!h->sort_auto || !h->slow_count) return;
    if (!h->slow_host) {
        if (cudaMallocHost((void **)&h->slow_host, 8 * sizeof(unsigned long long)) != cudaSuccess) { cudaGetLastError(); h->sort_auto = false; return; }
        for (int i = 0; i < 8; ++i) { h->slow_host[i] = 0; h->probe_step[i] = -1; }
    }
    const int64_t s = h->steps;
    if (s >= POLY_RUNAHEAD) {
        cudaEvent_t e = h->run_ev[(s - POLY_RUNAHEAD) & 7];
        if (e) cudaEventSynchronize(e);
        const int idx = (int)((s - POLY_RUNAHEAD) & 7);
        if (h->probe_step[idx] == s - POLY_RUNAHEAD) {
            const unsigned long long now = h->slow_host[idx];
            if (h->probe_have_prev && h->fs_enabled && h->sset_n == 0) {
                // the re-sort is fused into a step's passes and nearly free (picgolf: step_fixedpoint), so the interval may go down to
                // every step: never let the order get as old again as it was in a step that flushed a lot; grow (by half, up to
                // 64) only when the oldest order the interval allows was still quiet, and not for 32 steps after a cut
                const double frac = (double)(now - h->slow_seen) / (double)h->cfg.P; // flushes per particle during step s - RUNAHEAD - 1
                const int age = h->probe_age[(s - POLY_RUNAHEAD - (h->comm ? 2 : 1)) & 7];
                const double warps = (double)h->nblocks_poly * (CP_THREADS / 32) * (double)h->nranks;
                const double expect = poly_expected_flushes((double)h->cfg.N, CP_NSUB, h->nranks, warps, (double)h->cfg.P, h->det, h->sublg);
                PolySortPolicy pol{h->sort_every, h->grow_hold, h->poly_quiet};
                poly_sort_policy_probe(pol, frac, age, expect); // pg_sort_policy.h
                h->sort_every = pol.sort_every; h->grow_hold = pol.grow_hold; h->poly_quiet = pol.quiet;
            } else if (h->probe_have_prev) {
                const double frac = (double)(now - h->slow_seen) / (double)h->cfg.P; // flushes per particle during one step
                // expected in sorted order: the warp that streams a (cell, sign v) group walks through its CP_NSUB polynomial
                // intervals, and each of its 32 lanes flushes once per interval (plus once per warp range); ~4 passes per step
                const double warps = (double)h->nblocks_poly * (CP_THREADS / 32) * (double)h->nranks;
                const double expect = poly_expected_flushes((double)h->cfg.N, CP_NSUB, h->nranks, warps, (double)h->cfg.P, h->det, h->sublg);
                h->poly_quiet = frac < 5e-5 + 1.5 * expect;
                if (frac > 1e-3 + 3.0 * expect && h->since_sort >= 2 + POLY_RUNAHEAD) {
                    h->force_sort = true;
                    h->sort_every = (int)std::max<int64_t>(2, h->since_sort - POLY_RUNAHEAD);
                }
            }
            h->slow_seen = now;
            h->probe_have_prev = true;
        } else {
            h->probe_have_prev = false;
        }
    }
    // several GPUs: the sum over the ranks as the first solve of the PREVIOUS step saw it, i.e. everything up to the end of the step before
    // that one -- whole steps, like the single-GPU counter, one step later (the running sum flush_global is cut off in the middle of a step:
    // a step's final pass, where a stale order shows, would be billed to the step after it)
    const void *src = h->comm ? (const void *)&h->ctrl->flush_step : (const void *)h->slow_count;
    cudaMemcpyAsync(&h->slow_host[s & 7], src, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
    h->probe_step[s & 7] = s;
    h->probe_age[s & 7] = (int)std::min<int64_t>(h->since_sort, 1 << 20);
}

// Counting sort of the step-start state (xb[par], vb[par]) by cell into the other ping-pong buffers.
static int sort_particles_1d(picgolf_handle h)
{
    if (!h->poly) adapt_sort_interval(h);
    else if (h->sort_auto && !h->force_sort && h->poly_quiet && !(h->fs_enabled && h->sset_n == 0)) h->sort_every = std::min(64, h->sort_every + std::max(1, h->sort_every / 2));
    h->force_sort = false;
    const int sp = h->timer.begin(ST_SORT, h->stream);
    SortArgs a;
    memset(&a, 0, sizeof(a));
    a.in[0] = h->xb[h->par]; a.in[1] = h->vb[h->par];
    a.out[0] = h->xb[1 - h->par]; a.out[1] = h->vb[1 - h->par];
    a.pid_in = h->pid_valid ? h->pid[h->pidpar] : nullptr;
    a.pid_out = h->pid[1 - h->pidpar];
    a.bin_count = h->bin_count; a.bin_cursor = h->bin_cursor;
    a.P = h->count; a.narr = 2; a.nbins = h->nbins; a.mode = 0; a.N = (int)h->cfg.N; a.NY = 1; a.tshift = 0;
    a.vsplit = h->poly ? 1 : 0; a.sublg = h->sublg;
    int gh = (int)std::max<int64_t>(1, std::min<int64_t>((h->count + SORT_THREADS - 1) / SORT_THREADS, (int64_t)h->sms * 8));
    const int gs = h->poly ? 0 : sort_scatter_grid<2>(h->sms, h->count, h->nbins);
    if (h->poly) { // too many bins for shared-memory tables: warp-aggregated global atomics
        sort_hist_match_kernel<<<h->sms * 8, SORT_THREADS, 0, h->stream>>>(a);
        sort_scan_kernel<<<1, 1024, 0, h->stream>>>(h->bin_count, h->bin_cursor, nullptr, h->nbins);
        sort_scatter_match_kernel<2><<<h->sms * 8, SORT_THREADS, 0, h->stream>>>(a);
    } else {
        sort_hist_kernel<<<gh, SORT_THREADS, (size_t)h->nbins * 4, h->stream>>>(a);
        sort_scan_kernel<<<1, 1024, 0, h->stream>>>(h->bin_count, h->bin_cursor, nullptr, h->nbins);
        sort_scatter_kernel<2><<<gs, SORT_THREADS, (size_t)h->nbins * 8, h->stream>>>(a);
    }
    h->launches += 3;
    h->timer.end(sp, h->stream);
    h->par ^= 1; h->pidpar ^= 1; h->pid_valid = true; h->since_sort = 0; h->sorts++;
    return 0;
}

static FPArgs fp_args(picgolf_handle h)
{
    const picgolf_config &c = h->cfg;
    FPArgs a;
    a.X = h->xb[h->par]; a.V = h->vb[h->par]; a.v = h->vb[1 - h->par]; a.xout = h->xb[1 - h->par];
    a.E = h->E; a.rho = h->rho_fx; a.rho_next = h->rho_next; a.partials = h->partials; a.ctrl = h->ctrl;
    a.P = h->count; a.dt = c.dt; a.fx_scale = h->fx_scale; a.N = (int)c.N; a.k = 0;
    a.slow_count = h->slow_count; a.K = h->K; a.G = h->Gpoly; a.Mg = h->Mg;
    a.dN = (double)c.N * (CP_NSUB / 2); // polynomial passes: y = (x+X)*dN = c*N*CP_NSUB
    a.fs_hist = nullptr; a.fs_cursor = nullptr; a.fs_vout = nullptr; a.fs_pid_in = nullptr; a.fs_pid_out = nullptr;
    a.fs_scale = (double)((int64_t)c.N << h->sublg); a.fs_hs = c.dt / 2 * a.fs_scale;
    a.fs_magic = CP_MAGIC + (double)(1 << h->sublg) / 2; a.fs_sublg = h->sublg;
    if (h->fs_now) {
        a.fs_hist = h->bin_count; a.fs_cursor = h->bin_cursor; a.fs_vout = h->vspare;
        a.fs_pid_in = h->pid[h->pidpar]; a.fs_pid_out = h->pid[1 - h->pidpar];
    }
    return a;
}

static bool poly_now(picgolf_handle h) { return h->use_sorted_now && h->poly; }

// Pass 0 of the first step after the particles were set: later steps inherit the deposit fused into the previous final pass.
static int enqueue_first_pass(picgolf_handle h)
{
    FPArgs a = fp_args(h);
    const int sp = h->timer.begin(ST_PARTICLES, h->stream);
    if (poly_now(h) && h->det) fp_pass_poly<true, true><<<h->nblocks_poly, CP_THREADS, h->smem_poly, h->stream>>>(a);
    else if (poly_now(h)) fp_pass_poly<true, false><<<h->nblocks_poly, CP_THREADS, h->smem_poly, h->stream>>>(a);
    else if (h->use_sorted_now) fp_pass_sorted<true, SORTED_NP><<<h->nblocks_sorted, PG_THREADS, h->smem_sorted, h->stream>>>(a);
    else fp_pass_atomic<true><<<h->nblocks, PG_THREADS, h->smem_pass, h->stream>>>(a);
    h->timer.end(sp, h->stream);
    h->launches++;
    return 0;
}

// One sweep: [moments -> rho] [sum over ranks] solve [E -> per-cell gather polynomials (+ clear the moments)] particle pass.
// k >= 1: fixed schedule, the sweep index is a kernel argument and a sweep after convergence is a predicated no-op.
// k < 0:  body of the device-driven loop; the kernels read the sweep index from Ctrl and the solve steers the WHILE node `cond`.
static int enqueue_sweep(picgolf_handle h, int k, cudaGraphConditionalHandle cond)
{
    const picgolf_config &c = h->cfg;
    const int N = (int)c.N;
    FPArgs a = fp_args(h);
    a.k = k;
    const bool poly = poly_now(h);
    if (poly) { // cell-polynomial form (pg_kernels_poly.cuh)
        Mom2RhoArgs m;
        m.Mg = h->Mg; m.rho = h->rho_fx; m.ctrl = h->ctrl; m.fx_scale = h->fx_scale; m.fx_inv = h->fx_inv; m.N = N;
        m.flush_src = (h->comm && !h->peer_ok) ? h->slow_count : nullptr;
        m.det = h->det ? 1 : 0;
        const int sp = h->timer.begin(ST_SOLVE, h->stream);
        mom2rho_kernel<<<(N + CPM_CELLS - 1) / CPM_CELLS, 32 * CP_NSUB, CPM_SMEM, h->stream>>>(m);
        h->timer.end(sp, h->stream);
        h->launches++;
    }
    PG_TRY(allreduce_grid(h));
    PG_TRY(launch_solve1d(h, k, false, cond));
    if (poly) {
        GPolyArgs g;
        g.E = h->E; g.G = h->Gpoly; g.Mg = h->Mg; g.ctrl = h->ctrl; g.N = N; g.k = k; g.det = h->det ? 1 : 0;
        const int sp = h->timer.begin(ST_SOLVE, h->stream);
        gpoly_kernel<<<dim3((N + 127) / 128, CP_NSUB), 128, 0, h->stream>>>(g);
        h->timer.end(sp, h->stream);
        h->launches++;
    }
    if (poly && h->fs_now) { // fused re-sort: the bin counts of the previous pass -> slot cursors if this sweep is final, else cleared
        FsScanArgs f;
        f.hist = h->bin_count; f.cursor = h->bin_cursor; f.sync = h->fs_sync; f.ctrl = h->ctrl; f.nbins = h->nbins; f.k = k;
        const int sp = h->timer.begin(ST_SORT, h->stream);
        cp_fs_scan_kernel<<<(h->nbins + FS_CHUNK - 1) / FS_CHUNK, 1024, 0, h->stream>>>(f);
        h->timer.end(sp, h->stream);
        h->launches++;
    }
    const int sp = h->timer.begin(ST_PARTICLES, h->stream);
    if (poly && h->det) fp_pass_poly<false, true><<<h->nblocks_poly, CP_THREADS, h->smem_poly, h->stream>>>(a);
    else if (poly) fp_pass_poly<false, false><<<h->nblocks_poly, CP_THREADS, h->smem_poly, h->stream>>>(a);
    else if (h->use_sorted_now) fp_pass_sorted<false, SORTED_NP><<<h->nblocks_sorted, PG_THREADS, h->smem_sorted, h->stream>>>(a);
    else fp_pass_atomic<false><<<h->nblocks, PG_THREADS, h->smem_pass, h->stream>>>(a);
    h->timer.end(sp, h->stream);
    h->launches++;
    return 0;
}

// Fixed schedule: the host enqueues all max_sweeps sweeps; those after convergence are ~3 us predicated no-ops.  Used when
// the stage timers are on (their events cannot live inside a WHILE node) and when the grid is summed with NCCL.
static int enqueue_fixed_schedule(picgolf_handle h)
{
    for (int k = 1; k <= h->cfg.max_sweeps; ++k) PG_TRY(enqueue_sweep(h, k, 0));
    PG_TRY(launch_step_end(h, true));
    return 0;
}

// Device-driven loop: build (once per slot) the graph  WHILE(cond){ sweep } -> step_end.  cond starts at 1 on every launch
// (cudaGraphCondAssignDefault); the solve kernel clears it when isapprox(F,E) holds or k = max_sweeps, the rest of that
// iteration (the finalising particle pass) still runs, then the loop exits.
static int build_loop_graph(picgolf_handle h, cudaGraphExec_t *exec)
{
    if (!h->cap_stream) PG_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    PG_CUDA(cudaGraphCreate(&g, 0));
    struct Guard { cudaGraph_t g; ~Guard() { if (g) cudaGraphDestroy(g); } } guard{g};
    cudaGraphConditionalHandle cond;
    PG_CUDA(cudaGraphConditionalHandleCreate(&cond, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.conditional.handle = cond; cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
    cudaGraphNode_t wnode;
    PG_CUDA(cudaGraphAddNode(&wnode, g, nullptr, 0, &cp));
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    cudaStream_t run = h->stream;
    const int64_t l0 = h->launches;
    // the sweep, captured into the body of the WHILE node
    PG_CUDA(cudaStreamBeginCaptureToGraph(h->cap_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    h->stream = h->cap_stream;
    int rc = enqueue_sweep(h, -1, cond);
    h->stream = run;
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, nullptr);
    h->loop_launches_per_sweep = (int)(h->launches - l0) - (h->fs_now ? 1 : 0); // the bin scan of a re-sorting step is counted on the device (Ctrl.loop_fs_sweeps)
    h->launches = l0;
    if (rc != 0) return rc;
    PG_CUDA(e);
    // step_end, after the loop
    PG_CUDA(cudaStreamBeginCaptureToGraph(h->cap_stream, g, &wnode, nullptr, 1, cudaStreamCaptureModeThreadLocal));
    h->stream = h->cap_stream;
    rc = launch_step_end(h, true);
    h->stream = run;
    e = cudaStreamEndCapture(h->cap_stream, nullptr);
    h->launches = l0;
    if (rc != 0) return rc;
    PG_CUDA(e);
    PG_CUDA(cudaGraphInstantiate(exec, g, 0));
    return 0;
}

static int enqueue_simpson_step(picgolf_handle h)
{
    const picgolf_config &c = h->cfg;
    const int N = (int)c.N;
    SPArgs a;
    a.X = h->xb[h->par]; a.V = h->vb[h->par]; a.v = h->vb[1 - h->par]; a.xout = h->xb[1 - h->par];
    a.E = h->E; a.rho = h->rho_fx; a.partials = h->partials; a.ctrl = h->ctrl;
    a.P = h->count; a.dt = c.dt; a.fx_scale = h->fx_scale; a.N = N; a.k = 0;
    int sp = h->timer.begin(ST_PARTICLES, h->stream);
    const bool area = h->cfg.scheme == PICGOLF_AREA_SIMPSON13;
    if (area) sp_pass0<1><<<h->nblocks, PG_THREADS, (size_t)N * 8, h->stream>>>(a);
    else sp_pass0<0><<<h->nblocks, PG_THREADS, (size_t)N * 8, h->stream>>>(a);
    h->timer.end(sp, h->stream);
    h->launches++;
    PG_TRY(allreduce_grid(h, 0, 1));
    PG_TRY(launch_solve1d(h, 0, true)); // E[1,:] = solve(rho(X,X))
    sp = h->timer.begin(ST_PARTICLES, h->stream);
    if (area) sp_pass1<1><<<h->nblocks, PG_THREADS, h->smem_sp1, h->stream>>>(a);
    else sp_pass1<0><<<h->nblocks, PG_THREADS, h->smem_sp1, h->stream>>>(a);
    h->timer.end(sp, h->stream);
    h->launches++;
    SolveSPArgs s;
    s.rho_fx = h->rho_fx; s.rho_last = h->rho_last; s.E = h->E; s.tw = h->tw; s.ctrl = h->ctrl;
    s.w = c.w; s.fx_inv = h->fx_inv; s.rtol = c.rtol; s.atol = c.atol; s.N = N; s.lg = ilog2(c.N); s.max_sweeps = c.max_sweeps;
    const int threads = (int)std::min<int64_t>(1024, std::max<int64_t>(32, c.N / 2));
    for (int k = 1; k <= c.max_sweeps; ++k) {
        PG_TRY(allreduce_grid(h, 1, 2));
        s.k = k;
        sp = h->timer.begin(ST_SOLVE, h->stream);
        solve_simpson23_kernel<<<2, threads, h->smem_pass, h->stream>>>(s);
        h->timer.end(sp, h->stream);
        a.k = k;
        sp = h->timer.begin(ST_PARTICLES, h->stream);
        if (area) sp_passk<1><<<h->nblocks, PG_THREADS, h->smem_spk, h->stream>>>(a);
        else sp_passk<0><<<h->nblocks, PG_THREADS, h->smem_spk, h->stream>>>(a);
        h->timer.end(sp, h->stream);
        h->launches += 2;
    }
    PG_TRY(launch_step_end(h, true));
    return 0;
}

// Simpson-1/3 schemes: one step = 2*max_sweeps + 4 launches, most of them predicated no-ops.  For small problems the
// step is launch-bound, so the whole sequence is captured once per ping-pong parity into a CUDA graph and
// replayed (single GPU, stage timers off; NCCL calls and timing events stay out of graphs).
static int step_simpson(picgolf_handle h)
{
    const bool use_graph = !h->comm && !h->timer.enabled && h->count <= (1 << 22) && !h->graph_failed;
    if (!use_graph) {
        PG_TRY(enqueue_simpson_step(h));
    } else {
        const int slot = h->par + 2 * (h->have_deposit ? 1 : 0);
        cudaGraphExec_t &exec = h->step_graph[slot];
        if (!exec) {
            const int64_t l0 = h->launches;
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
            int rc = 0;
            if (e == cudaSuccess) {
                rc = enqueue_simpson_step(h);
                e = cudaStreamEndCapture(h->stream, &g);
            }
            if (e == cudaSuccess && rc == 0) e = cudaGraphInstantiate(&exec, g, 0);
            if (g) cudaGraphDestroy(g);
            h->graph_launches[slot] = h->launches - l0;
            h->launches = l0;
            if (e != cudaSuccess || rc != 0 || !exec) { // fall back to plain launches for good
                cudaGetLastError();
                exec = nullptr;
                h->graph_failed = true;
                PG_TRY(enqueue_simpson_step(h));
                h->par ^= 1; h->since_sort++;
                return 0;
            }
        }
        PG_CUDA(cudaGraphLaunch(exec, h->stream));
        h->launches += h->graph_launches[slot];
    }
    h->par ^= 1;
    h->since_sort++;
    return 0;
}

// One step of the Gaussian fixed point (GaussianFixedPoint.jl:7-10).  Default: the device-driven loop (build_loop_graph) --
// S sweeps cost S x {[mom2rho] [publish] solve [gpoly] pass} + step_end launches and nothing else.  The fixed schedule
// (all max_sweeps sweeps enqueued, the surplus predicated off) remains for stage timing and for the NCCL reduction.
static int step_fixedpoint(picgolf_handle h)
{
    // Lazy first sort: the first step after the particles were (re)set runs on the any-order kernels, the
    // cell sort happens before the second one.  A caller that exchanges the whole state every step (bench.py's
    // e2e arm) then never pays for a from-scratch sort + unsort that a single step cannot amortise.
    h->use_sorted_now = h->sorted && (h->pid_valid || h->steps > 0);
    if (h->use_sorted_now && h->poly && h->pid_valid) probe_poly_flushes(h);
    // polynomial passes on arrays that have been sorted once: the re-sort rides along with the passes of the LAST step the
    // current order is allowed to serve (pg_kernels_poly.cuh) -- the stand-alone sort runs before the first step of a new order
    const bool can_fuse = h->use_sorted_now && h->poly && h->pid_valid && h->fs_enabled && h->sset_n == 0 && !h->simpson;
    const bool due = h->use_sorted_now && (!h->pid_valid || h->force_sort || h->since_sort >= h->sort_every - (can_fuse ? 1 : 0));
    h->fs_now = due && can_fuse;
    if (due && !h->fs_now) PG_TRY(sort_particles_1d(h));
    if (h->simpson) return step_simpson(h);
    h->pass_blocks = poly_now(h) ? h->nblocks_poly : h->use_sorted_now ? h->nblocks_sorted : h->nblocks;
    if (!h->have_deposit) PG_TRY(enqueue_first_pass(h));
    const bool loop = !h->loop_off && !h->loop_failed && !h->timer.enabled && (!h->comm || h->peer_ok);
    bool done = false;
    if (loop) {
        const std::array<uintptr_t, 8> slot = {(uintptr_t)h->xb[h->par], (uintptr_t)h->vb[h->par], (uintptr_t)h->vb[1 - h->par], (uintptr_t)h->xb[1 - h->par],
                                               h->fs_now ? (uintptr_t)h->vspare : 0, (uintptr_t)h->rho_fx,
                                               (uintptr_t)((h->use_sorted_now ? 1 : 0) + (h->fs_now ? 2 + 4 * h->pidpar : 0)), 0};
        cudaGraphExec_t &exec = h->loop_graph[slot];
        if (!exec && build_loop_graph(h, &exec) != 0) { // no conditional nodes on this driver: fixed schedule for good
            cudaGetLastError();
            cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
            if (h->cap_stream && cudaStreamIsCapturing(h->cap_stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
                cudaGraph_t junk = nullptr;
                cudaStreamEndCapture(h->cap_stream, &junk);
                cudaGetLastError();
            }
            exec = nullptr;
            h->loop_failed = true;
        }
        if (exec) {
            PG_CUDA(cudaGraphLaunch(exec, h->stream));
            h->launches += 1; // step_end; the sweeps are counted on the device (Ctrl.loop_sweeps, picgolf_launch_count)
            h->loop_steps++;
            done = true;
        }
    }
    if (!done) PG_TRY(enqueue_fixed_schedule(h));
    if (h->fs_now) { // the final pass wrote x, v and the ids to their slots: v sits in the third buffer, the work buffer becomes the spare
        std::swap(h->vb[1 - h->par], h->vspare);
        h->pidpar ^= 1; h->sorts++; h->fused_sorts++; h->force_sort = false;
        h->since_sort = -1;
        h->fs_now = false;
    }
    h->par ^= 1;
    h->since_sort++;
    // the final pass deposited the next step's first rho (atomic / sorted kernels: into rho_next; polynomial: the moment grid)
    h->have_deposit = true;
    std::swap(h->rho_fx, h->rho_next);
    if (h->poly && h->sort_auto) { // end-of-step marker for probe_poly_flushes (h->steps is incremented by the caller)
        cudaEvent_t &e = h->run_ev[h->steps & 7];
        if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        if (e) cudaEventRecord(e, h->stream);
    }
    return 0;
}

static int lf_launch(picgolf_handle h, int do_kick, int do_deposit, int do_predrift = 0)
{
    const picgolf_config &c = h->cfg;
    LFArgs a;
    a.do_predrift = do_predrift;
    a.x = h->xb[0]; a.v = h->vb[0]; a.E = h->E; a.rho = h->rho_fx; a.partials = h->partials;
    a.P = h->count; a.dt = c.dt; a.fx_scale = h->fx_scale; a.N = (int)c.N; a.do_kick = do_kick; a.do_deposit = do_deposit;
    a.pow2 = h->dft ? 0 : 1;
    const int sp5_ = h->timer.begin(ST_PARTICLES, h->stream);
    const bool edge = do_predrift || do_deposit == 2; // first / last pass of a call: its own instantiation (lf_particle)
    if (h->ngp_tma) {
        if (edge) lf_pass_ngp_tma<true><<<h->nblocks, LF_TMA_THREADS, h->smem_lf, h->stream>>>(a);
        else lf_pass_ngp_tma<false><<<h->nblocks, LF_TMA_THREADS, h->smem_lf, h->stream>>>(a);
    } else if (h->ngp) {
        if (edge) lf_pass<0, true><<<h->nblocks, PG_THREADS, h->smem_lf, h->stream>>>(a);
        else lf_pass<0, false><<<h->nblocks, PG_THREADS, h->smem_lf, h->stream>>>(a);
    } else {
        if (edge) lf_pass<1, true><<<h->nblocks, PG_THREADS, h->smem_lf, h->stream>>>(a);
        else lf_pass<1, false><<<h->nblocks, PG_THREADS, h->smem_lf, h->stream>>>(a);
    }
    h->timer.end(sp5_, h->stream);
    h->launches++;
    return 0;
}

static int b1d2v_launch(picgolf_handle h, int do_push, int do_deposit)
{
    const picgolf_config &c = h->cfg;
    B1D2VArgs a;
    a.x = h->xb[0]; a.vx = h->vb[0]; a.vy = h->vy1; a.E = h->E; a.rho = h->rho_fx; a.partials = h->partials;
    a.P = h->count; a.dt = c.dt; a.B0 = c.B0; a.fx_scale = h->fx_scale;
    a.M = c.mass_ratio; a.first = h->first; a.Psp = c.scheme == PICGOLF_GAUSS_BORIS_1D2V2S ? c.P : -1;
    a.N = (int)c.N; a.do_push = do_push; a.do_deposit = do_deposit;
    const int sp = h->timer.begin(ST_PARTICLES, h->stream);
    b1d2v_pass<<<h->nblocks, PG_THREADS, h->smem_b1, h->stream>>>(a);
    h->timer.end(sp, h->stream);
    h->launches++;
    return 0;
}

static int launch_solve2d(picgolf_handle h)
{
    const picgolf_config &c = h->cfg;
    Solve2DArgs a;
    a.rho_in = nullptr; a.rho_fx = h->rho_fx; a.w = c.w; a.fx_inv = h->fx_inv;
    a.rho_last = h->rho_last; a.Z = h->Z; a.E2 = h->E2; a.twx = h->tw; a.twy = h->twy;
    a.partials = h->epartials; a.NX = (int)c.N; a.NY = (int)c.NY; a.lgx = ilog2(c.N); a.lgy = ilog2(c.NY);
    const int NX = a.NX, NY = a.NY;
    const int sp6_ = h->timer.begin(ST_SOLVE, h->stream);
    solve2d_rows_fwd<<<NY / ROWS_PER_BLOCK, 512, (size_t)2 * ROWS_PER_BLOCK * NX * 8, h->stream>>>(a);
    solve2d_cols<<<NX / COLS_PER_BLOCK, 512, (size_t)2 * COLS_PER_BLOCK * (NY + 1) * 8, h->stream>>>(a);
    solve2d_rows_inv<<<NY / ROWS_PER_BLOCK, 512, (size_t)(2 * ROWS_PER_BLOCK * NX + 32) * 8, h->stream>>>(a);
    h->timer.end(sp6_, h->stream);
    h->launches += 3;
    return 0;
}

// Counting sort of the 2D particle arrays by 16x16-cell tile + the per-tile work list.
static int sort_particles_2d(picgolf_handle h)
{
    adapt_sort_interval(h);
    const int sp = h->timer.begin(ST_SORT, h->stream);
    SortArgs a;
    memset(&a, 0, sizeof(a));
    for (int q = 0; q < 5; ++q) { a.in[q] = h->p2[h->par][q]; a.out[q] = h->p2[1 - h->par][q]; }
    a.pid_in = h->pid_valid ? h->pid[h->pidpar] : nullptr;
    a.pid_out = h->pid[1 - h->pidpar];
    a.bin_count = h->bin_count; a.bin_cursor = h->bin_cursor; a.bin_start = h->bin_start;
    a.P = h->count; a.narr = 5; a.nbins = h->nbins; a.mode = 1; a.N = (int)h->cfg.N; a.NY = (int)h->cfg.NY; a.tshift = T2_SHIFT;
    int gh = (int)std::max<int64_t>(1, std::min<int64_t>((h->count + SORT_THREADS - 1) / SORT_THREADS, (int64_t)h->sms * 8));
    const int gs = sort_scatter_grid<5>(h->sms, h->count, h->nbins);
    sort_hist_kernel<<<gh, SORT_THREADS, (size_t)h->nbins * 4, h->stream>>>(a);
    sort_scan_kernel<<<1, 1024, 0, h->stream>>>(h->bin_count, h->bin_cursor, h->bin_start, h->nbins);
    sort_scatter_kernel<5><<<gs, SORT_THREADS, (size_t)h->nbins * 8, h->stream>>>(a);
    tile_worklist_kernel<<<1, 1024, 0, h->stream>>>(h->bin_start, h->bin_cursor, h->item_off, h->nbins);
    h->launches += 4;
    h->timer.end(sp, h->stream);
    h->par ^= 1; h->pidpar ^= 1; h->pid_valid = true; h->since_sort = 0; h->sorts++;
    return 0;
}

static int step_2d3v(picgolf_handle h)
{
    const picgolf_config &c = h->cfg;
    h->use_sorted_now = h->sorted && (h->pid_valid || h->steps > 0); // lazy first sort, see step_fixedpoint
    if (h->use_sorted_now && (!h->pid_valid || h->since_sort >= h->sort_every)) PG_TRY(sort_particles_2d(h));
    P2DArgs a;
    memset(&a, 0, sizeof(a));
    double **p = h->p2[h->par];
    a.x = p[0]; a.y = p[1]; a.vx = p[2]; a.vy = p[3]; a.vz = p[4]; a.E2 = h->E2; a.rho = h->rho_fx;
    a.partials = h->partials; a.P = h->count; a.dt = c.dt; a.fx_scale = h->fx_scale;
    a.t1 = c.B0 * c.dt / 2;                       // tvec[1]   Electrostatic2D3V.jl:32
    a.tscale = 2 / (1 + (a.t1 * a.t1 + 0.0 + 0.0)); // :33
    a.NX = (int)c.N; a.NY = (int)c.NY;
    const int sp7_ = h->timer.begin(ST_PARTICLES, h->stream);
    h->pass_blocks = h->use_sorted_now ? h->nblocks_sorted : h->nblocks;
    if (h->use_sorted_now) {
        a.tile_start = h->bin_start; a.tile_end = h->bin_cursor; a.item_off = h->item_off; a.slow_count = h->slow_count;
        a.fxw_scale = h->fxw_scale; a.fx_shift = h->fx_shift; a.ntx = std::max(1, a.NX >> T2_SHIFT); a.ntiles = h->nbins;
        if (h->k2d == 2) stream_kernel(h->k2d_a, h->k2d_b, h->k2d_c)<<<h->nblocks_sorted, h->k2d_c, h->smem_ring, h->stream>>>(a);
        else particles_2d3v_tiled<<<h->nblocks_sorted, PG_THREADS, 0, h->stream>>>(a);
    } else {
        particles_2d3v_kernel<<<h->nblocks, PG_THREADS, 0, h->stream>>>(a);
    }
    h->timer.end(sp7_, h->stream);
    h->launches++;
    PG_TRY(allreduce_grid(h));
    PG_TRY(launch_solve2d(h));
    bool record = ((h->steps + 1) % c.diag_every) == 0; // if t % NS == 0   :164
    if (record && h->snap[0]) { // Exs[:,:,ti] .= real.(Ex); Eys ...; phis ...   :171-173, ti = t / NS
        const int64_t ti = (h->steps + 1) / c.diag_every - 1;
        if (ti < h->T) {
            const size_t off = (size_t)ti * h->ncell;
            snapshot2d_kernel<<<1, 1024, 0, h->stream>>>(h->E2, h->rho_last, h->ncell, h->snap[0] + off, h->snap[1] + off, h->snap[2] + off);
            h->launches++;
        }
    }
    PG_TRY(launch_step_end(h, record));
    h->since_sort++;
    return 0;
}

PG_API int picgolf_step(picgolf_handle h, int64_t nsteps)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (nsteps < 0) return fail(PICGOLF_ERR_ARG, "nsteps < 0");
    if (!h->have_particles) return fail(PICGOLF_ERR_STATE, "set or initialise particles before stepping");
    if (h->nranks > 1 && !h->comm)
        return fail(PICGOLF_ERR_STATE, "handle is rank %d of %d but picgolf_comm_init was never called: its shard alone would deposit a wrong charge", h->rank, h->nranks);
    if (nsteps == 0) return 0;
    PG_TRY(use_device(h));
    const int sp8_ = h->timer.begin(ST_TOTAL, h->stream);
    int rc = 0;
    if (h->fixedpoint) {
        for (int64_t s = 0; s < nsteps && rc == 0; ++s) { rc = step_fixedpoint(h); h->steps++; }
    } else if (h->is2d) {
        for (int64_t s = 0; s < nsteps && rc == 0; ++s) { rc = step_2d3v(h); h->steps++; }
    } else if (h->b1d2v) {
        if (!h->have_deposit) rc = b1d2v_launch(h, 0, 1); // rho(x) of the first step   NGP1D2V.jl:40
        for (int64_t s = 0; s < nsteps && rc == 0; ++s) {
            rc = allreduce_grid(h);
            if (rc == 0) rc = launch_solve1d(h, 1);
            if (rc == 0) rc = b1d2v_launch(h, 1, 1);     // gather, boris, move, wrap + rho(x) of the next step
            bool record = ((h->steps + 1) % h->cfg.diag_every) == 0; // if mod(t, T/TO) == 0   :58
            if (rc == 0) rc = launch_step_end(h, record);
            h->steps++;
        }
        h->have_deposit = true;
    } else {
        // A call ends on full-step positions (what picgolf_get_particles hands out), but its last pass has already deposited the
        // charge of the NEXT step at the half-drifted positions (do_deposit = 2); the next call redoes that half drift in its first
        // pass (do_predrift) instead of spending a pass of its own on u() + deposit: K steps cost K passes, not K + 1 -- a driver
        // that records something after every step (picgolf_step(h, 1) in a loop) runs twice as fast.
        const bool pending = h->have_deposit;
        if (!pending) rc = lf_launch(h, 0, 1); // first step after the particles were set: u(); deposit
        for (int64_t s = 0; s < nsteps && rc == 0; ++s) {
            rc = allreduce_grid(h);
            if (rc == 0) rc = launch_solve1d(h, 1);
            if (rc == 0) rc = lf_launch(h, 1, s + 1 < nsteps ? 1 : 2, (s == 0 && pending) ? 1 : 0); // u(); kick; u(); deposit of the next step
            if (rc == 0) rc = launch_step_end(h, true);
            h->steps++;
        }
        if (rc == 0) h->have_deposit = true;
    }
    h->timer.end(sp8_, h->stream);
    if (rc != 0) return rc;
    PG_CUDA(cudaGetLastError());
    if (h->timer.enabled && h->timer.open.size() > 4096) h->timer.drain();
    return 0;
}

// ------------------------------------------------------------------------------------------
// streamed step: set_particles -> step(1) -> get_particles as ONE asynchronous call, pipelined over three buffer sets
// ------------------------------------------------------------------------------------------
static void install_set(picgolf_handle h, int s)
{
    if (!h->sset_n || s == h->sset_cur) return;
    double **b = h->sset[s].a;
    if (h->is2d) for (int q = 0; q < 5; ++q) h->p2[0][q] = b[q];
    else { h->xb[0] = b[0]; h->xb[1] = b[1]; h->vb[0] = b[2]; h->vb[1] = b[3]; }
    h->sset_cur = s;
}

static int build_stream_ring(picgolf_handle h)
{
    if (h->sset_n) return 0;
    const size_t n = (size_t)h->count;
    double **b0 = h->sset[0].a;
    int na;
    if (h->is2d) { na = 5; for (int q = 0; q < 5; ++q) b0[q] = h->p2[0][q]; }
    else { na = 4; b0[0] = h->xb[0]; b0[1] = h->xb[1]; b0[2] = h->vb[0]; b0[3] = h->vb[1]; }
    for (int s = 1; s < 3; ++s)
        for (int i = 0; i < na; ++i)
            if (b0[i]) PG_TRY(dalloc(&h->sset[s].a[i], n)); // leapfrog schemes have no xb[1] / vb[1]
    PG_CUDA(cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking));
    PG_CUDA(cudaStreamCreateWithFlags(&h->down_stream, cudaStreamNonBlocking));
    for (int s = 0; s < 3; ++s) {
        PG_CUDA(cudaEventCreateWithFlags(&h->ev_up[s], cudaEventDisableTiming));
        PG_CUDA(cudaEventCreateWithFlags(&h->ev_comp[s], cudaEventDisableTiming));
        PG_CUDA(cudaEventCreateWithFlags(&h->ev_down[s], cudaEventDisableTiming));
    }
    h->sset_n = na; h->sset_cur = 0;
    return 0;
}

__global__ void stream_reset_kernel(Ctrl *c) { c->final_k = -1; c->sweeps = 0; }

// in[]/out[]: x, v (1D1V) or x, y, vx, vy, vz (2D3V) host arrays of the local shard.
static int step_streamed_impl(picgolf_handle h, const double *const *in, double *const *out, int narr, int64_t count)
{
    if (count != h->count) return fail(PICGOLF_ERR_ARG, "count %lld != local shard %lld", (long long)count, (long long)h->count);
    if (h->nranks > 1 && !h->comm) return fail(PICGOLF_ERR_STATE, "sharded handle without a communicator (picgolf_comm_init)");
    PG_TRY(use_device(h));
    if (!h->sset_n) { // first streamed call: the handle may hold a running simulation -- finish it, then build the ring
        PG_CUDA(cudaStreamSynchronize(h->stream));
        PG_TRY(bank_loop_launches(h));
        PG_TRY(build_stream_ring(h));
    }
    const int s = (int)(h->stream_calls % 3);
    const size_t bytes = (size_t)count * sizeof(double);
    double **b = h->sset[s].a;
    // upload: wait until the previous occupant of this set has been downloaded (which implies its step has finished)
    if (h->sset_used[s]) PG_CUDA(cudaStreamWaitEvent(h->up_stream, h->ev_down[s], 0));
    if (h->is2d) for (int q = 0; q < 5; ++q) PG_CUDA(cudaMemcpyAsync(b[q], in[q], bytes, cudaMemcpyHostToDevice, h->up_stream));
    else { PG_CUDA(cudaMemcpyAsync(b[0], in[0], bytes, cudaMemcpyHostToDevice, h->up_stream)); PG_CUDA(cudaMemcpyAsync(b[2], in[1], bytes, cudaMemcpyHostToDevice, h->up_stream)); }
    PG_CUDA(cudaEventRecord(h->ev_up[s], h->up_stream));
    // step: the state of reset_run_state(), enqueued instead of synchronised; the diagnostics trace keeps growing
    PG_CUDA(cudaStreamWaitEvent(h->stream, h->ev_up[s], 0));
    install_set(h, s);
    if (h->rho_base[0]) { h->rho_fx = h->rho_base[0]; h->rho_next = h->rho_base[1]; }
    PG_CUDA(cudaMemsetAsync(h->rho_fx, 0, (size_t)h->grid_rows * h->ncell * sizeof(unsigned long long), h->stream));
    if (h->rho_next) PG_CUDA(cudaMemsetAsync(h->rho_next, 0, ((size_t)h->ncell + 1) * sizeof(unsigned long long), h->stream));
    if (h->is2d) PG_CUDA(cudaMemsetAsync(h->E2, 0, h->ncell * sizeof(double2), h->stream));
    else PG_CUDA(cudaMemsetAsync(h->E, 0, (size_t)h->grid_rows * h->ncell * sizeof(double), h->stream));
    stream_reset_kernel<<<1, 1, 0, h->stream>>>(h->ctrl);
    h->launches++;
    h->par = 0; h->steps = 0; h->have_particles = true; h->pid_valid = false; h->pidpar = 0; h->since_sort = 0; h->have_deposit = false;
    h->slow_pending = false; h->force_sort = false; h->poly_quiet = false; h->probe_have_prev = false;
    for (auto &ps : h->probe_step) ps = -1;
    int rc = 0;
    if (h->fixedpoint) rc = step_fixedpoint(h);
    else if (h->is2d) rc = step_2d3v(h);
    else {
        rc = lf_launch(h, 0, 1);
        if (rc == 0) rc = allreduce_grid(h);
        if (rc == 0) rc = launch_solve1d(h, 1);
        if (rc == 0) rc = lf_launch(h, 1, 0);
        if (rc == 0) rc = launch_step_end(h, true);
    }
    h->steps++;
    if (rc != 0) return rc;
    PG_CUDA(cudaGetLastError());
    PG_CUDA(cudaEventRecord(h->ev_comp[s], h->stream));
    // download of the stepped state (1D fixed point: the other half of the set's ping-pong pair)
    PG_CUDA(cudaStreamWaitEvent(h->down_stream, h->ev_comp[s], 0));
    if (h->is2d) for (int q = 0; q < 5; ++q) { if (out[q]) PG_CUDA(cudaMemcpyAsync(out[q], h->p2[h->par][q], bytes, cudaMemcpyDeviceToHost, h->down_stream)); }
    else {
        if (out[0]) PG_CUDA(cudaMemcpyAsync(out[0], h->xb[h->par], bytes, cudaMemcpyDeviceToHost, h->down_stream));
        if (out[1]) PG_CUDA(cudaMemcpyAsync(out[1], h->vb[h->par], bytes, cudaMemcpyDeviceToHost, h->down_stream));
    }
    PG_CUDA(cudaEventRecord(h->ev_down[s], h->down_stream));
    h->sset_used[s] = true;
    h->stream_calls++;
    (void)narr;
    return 0;
}

PG_API int picgolf_step_streamed(picgolf_handle h, const double *x_in, const double *v_in, double *x_out, double *v_out, int64_t count)
{
    if (!h || !x_in || !v_in) return fail(PICGOLF_ERR_ARG, "NULL argument");
    if (h->is2d) return fail(PICGOLF_ERR_ARG, "use picgolf_step_streamed_2d3v for the 2D3V scheme");
    if (h->b1d2v || h->simpson) return fail(PICGOLF_ERR_UNSUPPORTED, "the streamed step is built for the 1D1V leapfrog / fixed-point schemes and for 2D3V");
    const double *in[2] = {x_in, v_in};
    double *out[2] = {x_out, v_out};
    return step_streamed_impl(h, in, out, 2, count);
}

PG_API int picgolf_step_streamed_2d3v(picgolf_handle h, const double *const in[5], double *const out[5], int64_t count)
{
    if (!h || !in || !out) return fail(PICGOLF_ERR_ARG, "NULL argument");
    if (!h->is2d) return fail(PICGOLF_ERR_ARG, "handle is not a 2D3V scheme");
    for (int q = 0; q < 5; ++q) if (!in[q]) return fail(PICGOLF_ERR_ARG, "NULL input array %d", q);
    return step_streamed_impl(h, in, out, 5, count);
}

PG_API int picgolf_steps_done(picgolf_handle h, int64_t *steps)
{
    if (!h || !steps) return fail(PICGOLF_ERR_ARG, "NULL argument");
    *steps = h->steps;
    return 0;
}

// ------------------------------------------------------------------------------------------
// fields and diagnostics
// ------------------------------------------------------------------------------------------
PG_API int picgolf_get_fields(picgolf_handle h, double *rho, double *E)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (h->is2d) return fail(PICGOLF_ERR_ARG, "use picgolf_get_fields_2d for the 2D3V scheme");
    PG_TRY(use_device(h));
    const size_t b = (size_t)h->ncell * sizeof(double);
    if (rho) PG_CUDA(cudaMemcpyAsync(rho, h->rho_last, b, cudaMemcpyDeviceToHost, h->stream));
    if (E) PG_CUDA(cudaMemcpyAsync(E, h->E + (size_t)(h->grid_rows - 1) * h->ncell, b, cudaMemcpyDeviceToHost, h->stream)); // E[end,:]
    PG_CUDA(cudaStreamSynchronize(h->stream));
    return check_peer(h);
}

PG_API int picgolf_set_field(picgolf_handle h, const double *E)
{
    if (!h || !E) return fail(PICGOLF_ERR_ARG, "NULL argument");
    if (h->is2d) return fail(PICGOLF_ERR_ARG, "picgolf_set_field is 1D only");
    PG_TRY(use_device(h));
    PG_CUDA(cudaMemcpyAsync(h->E + (size_t)(h->grid_rows - 1) * h->ncell, E, (size_t)h->ncell * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    PG_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

PG_API int picgolf_get_fields_2d(picgolf_handle h, double *rho, double *Ex, double *Ey)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (!h->is2d) return fail(PICGOLF_ERR_ARG, "handle is not a 2D3V scheme");
    PG_TRY(use_device(h));
    const size_t b = (size_t)h->ncell * sizeof(double);
    if (rho) PG_CUDA(cudaMemcpyAsync(rho, h->rho_last, b, cudaMemcpyDeviceToHost, h->stream));
    if (Ex || Ey) {
        double *tmp = nullptr;
        PG_TRY(dalloc(&tmp, (size_t)2 * h->ncell));
        split_E2_kernel<<<(unsigned)((h->ncell + 255) / 256), 256, 0, h->stream>>>(h->E2, h->ncell, tmp, tmp + h->ncell);
        h->launches++;
        if (Ex) PG_CUDA(cudaMemcpyAsync(Ex, tmp, b, cudaMemcpyDeviceToHost, h->stream));
        if (Ey) PG_CUDA(cudaMemcpyAsync(Ey, tmp + h->ncell, b, cudaMemcpyDeviceToHost, h->stream));
        PG_CUDA(cudaStreamSynchronize(h->stream));
        cudaFree(tmp);
    }
    PG_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

struct DevTmp {
    void *p = nullptr;
    ~DevTmp() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { PG_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 8))); return 0; }
};

// Fetch raw rows to the host; particle-derived columns are summed over ranks (collective call).
static int fetch_raw(picgolf_handle h, std::vector<double> &raw, int64_t *rows)
{
    PG_TRY(use_device(h));
    PG_CUDA(cudaStreamSynchronize(h->stream));
    PG_TRY(check_peer(h));
    Ctrl c;
    PG_CUDA(cudaMemcpy(&c, h->ctrl, sizeof(c), cudaMemcpyDeviceToHost));
    *rows = c.rows;
    raw.resize((size_t)4 * h->T);
    PG_CUDA(cudaMemcpy(raw.data(), h->raw, raw.size() * sizeof(double), cudaMemcpyDeviceToHost));
    if (h->comm) {
        // columns 1,2 (and 3 in 2D) hold local-shard sums: all-reduce them out of place.
        const size_t ncol = h->npart == 3 ? 3 : 2, cnt = ncol * (size_t)h->T; // particle-derived columns (sweeps column stays local)
        DevTmp tmp;
        PG_TRY(tmp.alloc(cnt * sizeof(double)));
        PG_TRY(nccl::check(nccl::AllReduce(h->raw + h->T, tmp.p, cnt, nccl::Float64, nccl::Sum, h->comm, h->stream),
                           "ncclAllReduce(diagnostics)"));
        PG_CUDA(cudaStreamSynchronize(h->stream));
        PG_CUDA(cudaMemcpy(raw.data() + h->T, tmp.p, cnt * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return 0;
}

PG_API int picgolf_get_raw_diagnostics(picgolf_handle h, double *raw, int64_t ld, int64_t *rows_out)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    std::vector<double> r;
    int64_t rows = 0;
    PG_TRY(fetch_raw(h, r, &rows));
    if (rows_out) *rows_out = rows;
    if (raw) {
        if (ld < rows) return fail(PICGOLF_ERR_ARG, "ld %lld < rows %lld", (long long)ld, (long long)rows);
        for (int c = 0; c < 4; ++c)
            for (int64_t t = 0; t < rows; ++t) raw[c * ld + t] = r[c * h->T + t];
    }
    return 0;
}

PG_API int picgolf_get_diagnostics(picgolf_handle h, double *D, int64_t ld, int32_t *sweeps, int64_t *rows_out)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    std::vector<double> r;
    int64_t rows = 0;
    PG_TRY(fetch_raw(h, r, &rows));
    if (rows_out) *rows_out = rows;
    if (D && ld < rows) return fail(PICGOLF_ERR_ARG, "ld %lld < rows %lld", (long long)ld, (long long)rows);
    const picgolf_config &c = h->cfg;
    const int64_t T = h->T;
    for (int64_t t = 0; t < rows; ++t) {
        double se = r[t], s1 = r[T + t], s2 = r[2 * T + t], s3 = r[3 * T + t];
        if (h->b1d2v) {
            if (D) {
                // D[ti,1:2].+=(sum(abs2,E)/N, sum(vy^2+vx^2)*n0/P)./2; D[ti,3:5].+=sum.((D[ti,1:2],vx/P,vy/P));
                // D[ti,1:3].*=2/n0; ... D ./= T/TO        NGP1D2V.jl:59-61,64
                double d1 = (se / (double)c.N) / 2, d2 = (s1 * c.W / (double)c.P) / 2, d3 = d1 + d2, sc = 2 / c.W;
                double win = (double)c.diag_every;
                D[t] = d1 * sc / win; D[ld + t] = d2 * sc / win; D[2 * ld + t] = d3 * sc / win;
                D[3 * ld + t] = s2 / (double)c.P / win; D[4 * ld + t] = s3 / (double)c.P / win;
            }
            if (sweeps) sweeps[t] = 1;
        } else if (!h->is2d) {
            if (D) {
                // D[t,1:2].=(sum(E.^2)/N,sum(v.^2)*W/P)./2; D[t,3:4].=sum.((D[t,1:2],v/P)); D[t,1:3].*=2/W
                double d1 = (se / (double)c.N) / 2, d2 = (s1 * c.W / (double)c.P) / 2;
                double d3 = d1 + d2, sc = 2 / c.W;
                D[t] = d1 * sc; D[ld + t] = d2 * sc; D[2 * ld + t] = d3 * sc; D[3 * ld + t] = s2 / (double)c.P;
            }
            if (sweeps) sweeps[t] = (int32_t)s3;
        } else {
            if (D) {
                // K[ti,1]=mean(Ex^2+Ey^2); K2=sum((vx^2+vy^2)*w); K3=K1+K2; K4=sum(vx)/P; K5=sum(vy)/P   :166-170
                double k1 = se / (double)(c.N * c.NY), k2 = s1 * c.w;
                D[t] = k1; D[ld + t] = k2; D[2 * ld + t] = k1 + k2; D[3 * ld + t] = s2 / (double)c.P; D[4 * ld + t] = s3 / (double)c.P;
            }
            if (sweeps) sweeps[t] = 1;
        }
    }
    return 0;
}

PG_API int picgolf_get_field_history(picgolf_handle h, double *Es, int64_t max_cols, int64_t *cols_out)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (!h->b1d2v) return fail(PICGOLF_ERR_ARG, "field history is kept by the 1D2V scheme only");
    PG_TRY(use_device(h));
    PG_CUDA(cudaStreamSynchronize(h->stream));
    const picgolf_config &c = h->cfg;
    int64_t cols = std::min<int64_t>(h->T, (h->steps + c.diag_every - 1) / c.diag_every);
    if (cols_out) *cols_out = cols;
    if (Es) {
        if (max_cols < cols) return fail(PICGOLF_ERR_ARG, "max_cols %lld < %lld", (long long)max_cols, (long long)cols);
        PG_CUDA(cudaMemcpy(Es, h->hist, (size_t)cols * c.N * sizeof(double), cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < cols * c.N; ++i) Es[i] /= (double)c.diag_every; // Es ./= T/TO
    }
    return 0;
}

PG_API int picgolf_get_snapshots_2d(picgolf_handle h, int which, double *out, int64_t max_slices, int64_t *slices_out)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (!h->is2d) return fail(PICGOLF_ERR_ARG, "handle is not a 2D3V scheme");
    if (which < 0 || which > 2) return fail(PICGOLF_ERR_ARG, "which must be 0 (Exs), 1 (Eys) or 2 (phis)");
    if (!h->snap[0]) return fail(PICGOLF_ERR_STATE, "the handle was created with field_history = 0");
    PG_TRY(use_device(h));
    PG_CUDA(cudaStreamSynchronize(h->stream));
    const int64_t slices = std::min<int64_t>(h->T, h->steps / h->cfg.diag_every);
    if (slices_out) *slices_out = slices;
    if (out) {
        if (max_slices < slices) return fail(PICGOLF_ERR_ARG, "max_slices %lld < %lld", (long long)max_slices, (long long)slices);
        PG_CUDA(cudaMemcpy(out, h->snap[which], (size_t)slices * h->ncell * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return 0;
}

PG_API int picgolf_stage_timing(picgolf_handle h, int enable)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    PG_TRY(use_device(h));
    if (!enable) h->timer.drain();
    h->timer.enabled = enable != 0;
    return 0;
}

PG_API int picgolf_stage_times(picgolf_handle h, double ms[5], int reset)
{
    if (!h || !ms) return fail(PICGOLF_ERR_ARG, "NULL argument");
    PG_TRY(use_device(h));
    PG_CUDA(cudaStreamSynchronize(h->stream));
    h->timer.drain();
    for (int i = 0; i < 5; ++i) ms[i] = h->timer.ms[i];
    if (reset) for (int i = 0; i < 5; ++i) h->timer.ms[i] = 0;
    return 0;
}

PG_API int picgolf_launch_count(picgolf_handle h, int64_t *launches)
{
    if (!h || !launches) return fail(PICGOLF_ERR_ARG, "NULL argument");
    *launches = h->launches;
    if (h->loop_steps > 0) { // + the sweeps the device-driven loop has launched since the particles were set
        PG_TRY(use_device(h));
        PG_CUDA(cudaStreamSynchronize(h->stream));
        Ctrl c;
        PG_CUDA(cudaMemcpy(&c, h->ctrl, sizeof(c), cudaMemcpyDeviceToHost));
        *launches += (int64_t)c.loop_sweeps * h->loop_launches_per_sweep + (int64_t)c.loop_fs_sweeps;
    }
    return 0;
}

PG_API int picgolf_sort_stats(picgolf_handle h, int64_t *sorts, int64_t *slow_particles)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    PG_TRY(use_device(h));
    if (sorts) *sorts = h->sorts;
    if (slow_particles) {
        unsigned long long n = 0;
        if (h->slow_count) {
            PG_CUDA(cudaStreamSynchronize(h->stream));
            PG_CUDA(cudaMemcpy(&n, h->slow_count, sizeof(n), cudaMemcpyDeviceToHost));
        }
        *slow_particles = (int64_t)n;
    }
    return 0;
}

PG_API int picgolf_fused_sorts(picgolf_handle h, int64_t *n)
{
    if (!h || !n) return fail(PICGOLF_ERR_ARG, "null argument");
    *n = h->fused_sorts;
    return 0;
}

PG_API int picgolf_deposit_path(picgolf_handle h, int *mode)
{
    if (!h || !mode) return fail(PICGOLF_ERR_ARG, "NULL argument");
    *mode = h->poly ? PICGOLF_DEPOSIT_POLY : h->sorted ? PICGOLF_DEPOSIT_SORTED : PICGOLF_DEPOSIT_ATOMIC;
    return 0;
}

PG_API int picgolf_get_stream(picgolf_handle h, void **stream)
{
    if (!h || !stream) return fail(PICGOLF_ERR_ARG, "NULL argument");
    *stream = (void *)h->stream;
    return 0;
}

// ------------------------------------------------------------------------------------------
// multi-GPU
// ------------------------------------------------------------------------------------------
// Peer-memory reduction (pg_peer.cuh).  export: allocate this rank's published-grid buffer and return its 64-byte
// cudaIpc handle; connect: open the nranks handles (own slot: the local pointer).  Both after picgolf_comm_init
// (NCCL stays for the diagnostics sums, the 2D and Simpson grids, and as the fallback when this is not set up).
PG_API int picgolf_peer_export(picgolf_handle h, void *handle64)
{
    if (!h || !handle64) return fail(PICGOLF_ERR_ARG, "NULL argument");
    if (h->nranks > PEER_MAX) return fail(PICGOLF_ERR_UNSUPPORTED, "peer reduction supports up to %d ranks", PEER_MAX);
    PG_TRY(use_device(h));
    if (!h->peer_mine) {
        const size_t bytes = sizeof(PeerPub) + (size_t)2 * h->ncell * sizeof(unsigned long long);
        PG_CUDA(cudaMalloc((void **)&h->peer_mine, bytes));
        PG_CUDA(cudaMemset(h->peer_mine, 0, bytes));
        PG_TRY(dalloc(&h->peer_err, 1));
        PG_CUDA(cudaMemset(h->peer_err, 0, sizeof(int)));
    }
    cudaIpcMemHandle_t ipc;
    PG_CUDA(cudaIpcGetMemHandle(&ipc, h->peer_mine));
    static_assert(sizeof(ipc) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &ipc, 64);
    return 0;
}

PG_API int picgolf_peer_connect(picgolf_handle h, const void *handles, int nranks, int rank)
{
    if (!h || !handles) return fail(PICGOLF_ERR_ARG, "NULL argument");
    if (nranks != h->nranks || rank != h->rank) return fail(PICGOLF_ERR_ARG, "peer_connect (rank %d of %d) does not match the handle (rank %d of %d)", rank, nranks, h->rank, h->nranks);
    if (!h->peer_mine) return fail(PICGOLF_ERR_STATE, "call picgolf_peer_export first");
    if (!h->comm) return fail(PICGOLF_ERR_STATE, "call picgolf_comm_init first");
    PG_TRY(use_device(h));
    for (int q = 0; q < nranks; ++q) {
        if (q == rank) { h->peer_ptr[q] = h->peer_mine; continue; }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, (const char *)handles + (size_t)64 * q, 64);
        void *p = nullptr;
        PG_CUDA(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
        h->peer_ptr[q] = (PeerPub *)p;
    }
    h->peer_ok = true;
    // map every peer buffer now (cudaIpc maps lazily): the first access over NVLink must not land inside a timed sweep
    peer_touch_kernel<<<1, 1, 0, h->stream>>>(peer_args(h), (unsigned long long *)h->peer_err);
    h->launches++;
    PG_CUDA(cudaGetLastError());
    PG_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

PG_API int picgolf_peer_status(picgolf_handle h, int *enabled, int *timed_out)
{
    if (!h) return fail(PICGOLF_ERR_ARG, "NULL handle");
    if (enabled) *enabled = h->peer_ok ? 1 : 0;
    if (timed_out) {
        *timed_out = 0;
        if (h->peer_err) {
            PG_TRY(use_device(h));
            PG_CUDA(cudaStreamSynchronize(h->stream));
            PG_CUDA(cudaMemcpy(timed_out, h->peer_err, sizeof(int), cudaMemcpyDeviceToHost));
        }
    }
    return 0;
}

PG_API int picgolf_comm_unique_id(void *id128)
{
    if (!id128) return fail(PICGOLF_ERR_ARG, "NULL argument");
    PG_TRY(nccl::load());
    nccl::UniqueId id;
    PG_TRY(nccl::check(nccl::GetUniqueId(&id), "ncclGetUniqueId"));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

PG_API int picgolf_comm_init(picgolf_handle h, const void *id128, int nranks, int rank)
{
    if (!h || !id128) return fail(PICGOLF_ERR_ARG, "NULL argument");
    if (nranks != h->nranks || rank != h->rank)
        return fail(PICGOLF_ERR_ARG, "comm (rank %d of %d) does not match the handle's shard (rank %d of %d)", rank, nranks, h->rank, h->nranks);
    if (nranks == 1) return 0;
    PG_TRY(use_device(h));
    PG_TRY(nccl::load());
    nccl::UniqueId id;
    memcpy(&id, id128, sizeof(id));
    PG_TRY(nccl::check(nccl::CommInitRank(&h->comm, nranks, id, rank), "ncclCommInitRank"));
    return 0;
}

// ------------------------------------------------------------------------------------------
// stage-level entry points
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { PG_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 8))); return 0; }
    int upload(const void *src, size_t bytes) { PG_TRY(alloc(bytes)); PG_CUDA(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice)); return 0; }
    int zero(size_t bytes) { PG_TRY(alloc(bytes)); PG_CUDA(cudaMemset(p, 0, bytes)); return 0; }
    int download(void *dst, size_t bytes) { PG_CUDA(cudaMemcpy(dst, p, bytes, cudaMemcpyDeviceToHost)); return 0; }
    template <typename T> T *as() { return (T *)p; }
};

static int stage_ready()
{
    if (picgolf_device_count() < 1) return fail(PICGOLF_ERR_CUDA, "no CUDA device: libpicgolf has no CPU path");
    return 0;
}
static unsigned grid1(int64_t n) { return (unsigned)std::max<int64_t>(1, (n + 255) / 256); }
static int finish() { PG_CUDA(cudaGetLastError()); PG_CUDA(cudaDeviceSynchronize()); return 0; }
static int check_grid1d(int64_t N)
{
    if (N < 16 || N > 8192) return fail(PICGOLF_ERR_ARG, "N must be in 16..8192");
    if (!is_pow2(N)) return fail(PICGOLF_ERR_UNSUPPORTED, "N=%lld: only power-of-two grids are built", (long long)N);
    return 0;
}
// the NGP stages (deposit, solve) also take even grids that are not a power of two, like the NGP leapfrog itself
static int check_grid1d_ngp(int64_t N)
{
    if (N < 16 || N > 8192) return fail(PICGOLF_ERR_ARG, "N must be in 16..8192");
    if (!is_pow2(N) && (N & 1)) return fail(PICGOLF_ERR_UNSUPPORTED, "N=%lld: grids that are not a power of two must be even", (long long)N);
    return 0;
}

PG_API int picgolf_stage_ngp_index(const double *x, int64_t count, int64_t N, int32_t *idx1)
{
    if (!x || !idx1 || count < 0 || N < 1) return fail(PICGOLF_ERR_ARG, "bad argument");
    PG_TRY(stage_ready());
    DevBuf dx, di;
    PG_TRY(dx.upload(x, count * 8)); PG_TRY(di.alloc(count * 4));
    stage_ngp_index_kernel<<<grid1(count), 256>>>(dx.as<double>(), count, (int)N, di.as<int>());
    PG_TRY(finish());
    return di.download(idx1, count * 4);
}

PG_API int picgolf_stage_mod1(const double *x, int64_t count, double *out)
{
    if (!x || !out || count < 0) return fail(PICGOLF_ERR_ARG, "bad argument");
    PG_TRY(stage_ready());
    DevBuf dx, dout;
    PG_TRY(dx.upload(x, count * 8)); PG_TRY(dout.alloc(count * 8));
    stage_mod1_kernel<<<grid1(count), 256>>>(dx.as<double>(), count, dout.as<double>());
    PG_TRY(finish());
    return dout.download(out, count * 8);
}

PG_API int picgolf_stage_gauss_stencil(const double *c, int64_t count, int64_t N, int hw, int32_t *idx1, double *wt)
{
    if (!c || !idx1 || !wt || count < 0) return fail(PICGOLF_ERR_ARG, "bad argument");
    if (hw != 6 && hw != 7) return fail(PICGOLF_ERR_ARG, "half_width must be 6 or 7");
    PG_TRY(check_grid1d(N)); PG_TRY(stage_ready());
    const size_t nw = 2 * hw + 1;
    DevBuf dc, di, dw;
    PG_TRY(dc.upload(c, count * 8)); PG_TRY(di.alloc(count * nw * 4)); PG_TRY(dw.alloc(count * nw * 8));
    stage_gauss_stencil_kernel<<<grid1(count), 256>>>(dc.as<double>(), count, (int)N, hw, di.as<int>(), dw.as<double>());
    PG_TRY(finish());
    PG_TRY(di.download(idx1, count * nw * 4));
    return dw.download(wt, count * nw * 8);
}

PG_API int picgolf_stage_ngp_deposit(const double *x, int64_t count, int64_t N, double w, double *rho)
{
    if (!x || !rho || count < 0) return fail(PICGOLF_ERR_ARG, "bad argument");
    PG_TRY(check_grid1d_ngp(N)); PG_TRY(stage_ready());
    // Runs the production pass (deposit half only) and the production solve's count->rho conversion.
    DevBuf dx, dv, dcnt, dpart;
    std::vector<double> zeros((size_t)count, 0.0);
    PG_TRY(dx.upload(x, count * 8)); PG_TRY(dv.upload(zeros.data(), count * 8));
    PG_TRY(dcnt.zero(N * 8)); PG_TRY(dpart.zero(2 * 1024 * 8));
    size_t smem = lf_smem_bytes(0, (int)N);
    PG_TRY(set_smem(lf_pass<0, false>, smem));
    LFArgs a;
    a.x = dx.as<double>(); a.v = dv.as<double>(); a.E = nullptr; a.rho = dcnt.as<unsigned long long>();
    a.partials = dpart.as<double>(); a.P = count; a.dt = 0.0; a.fx_scale = 1.0; a.N = (int)N; a.do_kick = 0; a.do_deposit = 1;
    a.pow2 = is_pow2(N) ? 1 : 0;
    // with v = 0 and dt = 0 the half drift is x = mod(x + 0, 1): positions in [0,1) are unchanged
    lf_pass<0, false><<<(unsigned)std::min<int64_t>(grid1(count), 1024), PG_THREADS, smem>>>(a);
    PG_TRY(finish());
    std::vector<unsigned long long> cnt((size_t)N);
    PG_TRY(dcnt.download(cnt.data(), N * 8));
    for (int64_t n = 0; n < N; ++n) rho[n] = (double)cnt[n] * w; // same expression as solve1d_kernel
    return 0;
}

PG_API int picgolf_stage_gauss_deposit(const double *x, const double *y, int64_t count, int64_t N, int hw, double w,
                                       int mode, double *rho)
{
    if (!x || !y || !rho || count < 0) return fail(PICGOLF_ERR_ARG, "bad argument");
    if (hw != 6 && hw != 7) return fail(PICGOLF_ERR_ARG, "half_width must be 6 or 7");
    (void)mode;
    PG_TRY(check_grid1d(N)); PG_TRY(stage_ready());
    DevBuf dx, dy, dr;
    PG_TRY(dx.upload(x, count * 8)); PG_TRY(dy.upload(y, count * 8)); PG_TRY(dr.zero(N * 8));
    size_t smem = (size_t)N * 8;
    PG_TRY(set_smem(stage_gauss_deposit_kernel, smem));
    // same fixed-point format rule as picgolf_create
    int frac = std::max(8, std::min(60, 62 - ilog2(count + 1)));
    stage_gauss_deposit_kernel<<<(unsigned)std::min<int64_t>(grid1(count), 1024), PG_THREADS, smem>>>(
        dx.as<double>(), dy.as<double>(), count, (int)N, ldexp(1.0, frac), dr.as<unsigned long long>());
    PG_TRY(finish());
    std::vector<long long> fx((size_t)N);
    PG_TRY(dr.download(fx.data(), N * 8));
    for (int64_t n = 0; n < N; ++n) rho[n] = (double)fx[n] * ldexp(1.0, -frac) * w; // same expression as solve1d_kernel
    return 0;
}

PG_API int picgolf_stage_gauss_gather(const double *E, int64_t N, int hw, const double *c, int64_t count, double *out)
{
    if (!E || !c || !out || count < 0) return fail(PICGOLF_ERR_ARG, "bad argument");
    if (hw != 6 && hw != 7) return fail(PICGOLF_ERR_ARG, "half_width must be 6 or 7");
    PG_TRY(check_grid1d(N)); PG_TRY(stage_ready());
    DevBuf dE, dc, dout;
    PG_TRY(dE.upload(E, N * 8)); PG_TRY(dc.upload(c, count * 8)); PG_TRY(dout.alloc(count * 8));
    stage_gauss_gather_kernel<<<grid1(count), 256>>>(dE.as<double>(), (int)N, dc.as<double>(), count, dout.as<double>());
    PG_TRY(finish());
    return dout.download(out, count * 8);
}

PG_API int picgolf_stage_solve1d(const double *rho, int64_t N, double *E)
{
    if (!rho || !E) return fail(PICGOLF_ERR_ARG, "bad argument");
    PG_TRY(check_grid1d_ngp(N)); PG_TRY(stage_ready());
    DevBuf dr, dl, dE, dctrl;
    double2 *tw = nullptr;
    PG_TRY(dr.upload(rho, N * 8)); PG_TRY(dl.alloc(N * 8)); PG_TRY(dE.zero(N * 8)); PG_TRY(dctrl.zero(sizeof(Ctrl)));
    if (!is_pow2(N)) { // direct transforms (solve1d_dft_*)
        DevBuf dspec, dpart, darr;
        const int nb = dft_blocks((int)N);
        PG_TRY(dspec.alloc(N * 16)); PG_TRY(dpart.zero((size_t)nb * 8)); PG_TRY(darr.zero(8));
        double2 *dtw = nullptr;
        PG_TRY(make_dft_twiddles(&dtw, (int)N));
        struct FreeTw { double2 *p; ~FreeTw() { if (p) cudaFree(p); } } free_tw{dtw};
        SolveDftArgs d;
        memset(&d, 0, sizeof(d));
        d.tw = dtw;
        d.rho_in = dr.as<double>(); d.rho_last = dl.as<double>(); d.E = dE.as<double>(); d.spec = dspec.as<double2>();
        d.part = dpart.as<double>(); d.arrive = darr.as<unsigned int>(); d.ctrl = dctrl.as<Ctrl>(); d.w = 1.0; d.fx_inv = 1.0; d.N = (int)N;
        const size_t sm = dft_smem_bytes((int)N);
        PG_TRY(set_smem(solve1d_dft_fwd, sm)); PG_TRY(set_smem(solve1d_dft_inv, sm));
        solve1d_dft_fwd<<<nb, DFT_THREADS, sm>>>(d);
        solve1d_dft_inv<<<nb, DFT_THREADS, sm>>>(d);
        PG_TRY(finish());
        return dE.download(E, N * 8);
    }
    PG_TRY(make_twiddles(&tw, (int)N));
    Solve1DArgs a;
    memset(&a, 0, sizeof(a));
    a.rho_in = dr.as<double>(); a.rho_fx = nullptr; a.rho_last = dl.as<double>(); a.E = dE.as<double>(); a.tw = tw;
    a.ctrl = dctrl.as<Ctrl>(); a.w = 1.0; a.fx_inv = 1.0; a.rtol = 0; a.atol = 0; a.N = (int)N; a.lg = ilog2(N);
    a.fixedpoint = 0; a.k = 1; a.max_sweeps = 1; a.store_normE1 = 0; a.hist = nullptr;
    int rc = set_smem_solve1d((int)N);
    if (rc == 0) {
        launch_solve1d_kernel(a, 0);
        rc = finish();
    }
    if (rc == 0) rc = dE.download(E, N * 8);
    cudaFree(tw);
    return rc;
}

#ifdef PG_SOLVE_PROF
PG_API int picgolf_debug_solve_prof(long long *out8) { return cudaMemcpyFromSymbol(out8, g_solve_prof, 8 * sizeof(long long)) == cudaSuccess ? 0 : -3; }
#endif

PG_API int picgolf_stage_solve2d(const double *rho, int64_t NX, int64_t NY, double *Ex, double *Ey)
{
    if (!rho || !Ex || !Ey) return fail(PICGOLF_ERR_ARG, "bad argument");
    if (NX < 16 || NY < 16 || NX > 1024 || NY > 1024) return fail(PICGOLF_ERR_ARG, "2D grid must be 16..1024 per side");
    if (!is_pow2(NX) || !is_pow2(NY)) return fail(PICGOLF_ERR_UNSUPPORTED, "only power-of-two grids are built");
    PG_TRY(stage_ready());
    const size_t n = (size_t)NX * NY;
    DevBuf dr, dl, dZ, dE2, dp, dsplit;
    double2 *twx = nullptr, *twy = nullptr;
    PG_TRY(dr.upload(rho, n * 8)); PG_TRY(dl.alloc(n * 8)); PG_TRY(dZ.alloc(n * 16)); PG_TRY(dE2.alloc(n * 16));
    PG_TRY(dp.alloc((NY / ROWS_PER_BLOCK) * 8)); PG_TRY(dsplit.alloc(2 * n * 8));
    PG_TRY(make_twiddles(&twx, (int)NX));
    int rc = make_twiddles(&twy, (int)NY);
    Solve2DArgs a;
    a.rho_in = dr.as<double>(); a.rho_fx = nullptr; a.w = 1.0; a.fx_inv = 1.0;
    a.rho_last = dl.as<double>(); a.Z = dZ.as<double2>(); a.E2 = dE2.as<double2>();
    a.twx = twx; a.twy = twy; a.partials = dp.as<double>(); a.NX = (int)NX; a.NY = (int)NY; a.lgx = ilog2(NX); a.lgy = ilog2(NY);
    if (rc == 0) rc = set_smem(solve2d_rows_fwd, (size_t)2 * ROWS_PER_BLOCK * NX * 8);
    if (rc == 0) rc = set_smem(solve2d_rows_inv, (size_t)(2 * ROWS_PER_BLOCK * NX + 32) * 8);
    if (rc == 0) rc = set_smem(solve2d_cols, (size_t)2 * COLS_PER_BLOCK * (NY + 1) * 8);
    if (rc == 0) {
        solve2d_rows_fwd<<<(unsigned)(NY / ROWS_PER_BLOCK), 512, (size_t)2 * ROWS_PER_BLOCK * NX * 8>>>(a);
        solve2d_cols<<<(unsigned)(NX / COLS_PER_BLOCK), 512, (size_t)2 * COLS_PER_BLOCK * (NY + 1) * 8>>>(a);
        solve2d_rows_inv<<<(unsigned)(NY / ROWS_PER_BLOCK), 512, (size_t)(2 * ROWS_PER_BLOCK * NX + 32) * 8>>>(a);
        split_E2_kernel<<<grid1((int64_t)n), 256>>>(dE2.as<double2>(), (long long)n, dsplit.as<double>(), dsplit.as<double>() + n);
        rc = finish();
    }
    if (rc == 0) rc = dsplit.download(Ex, n * 8);
    if (rc == 0) PG_CUDA(cudaMemcpy(Ey, dsplit.as<double>() + n, n * 8, cudaMemcpyDeviceToHost));
    cudaFree(twx); cudaFree(twy);
    return rc;
}

PG_API int picgolf_stage_cic_deposit(const double *x, const double *y, int64_t count, int64_t NX, int64_t NY, double w,
                                     double *rho)
{
    if (!x || !y || !rho || count < 0 || NX < 1 || NY < 1) return fail(PICGOLF_ERR_ARG, "bad argument");
    PG_TRY(stage_ready());
    DevBuf dx, dy, dr;
    PG_TRY(dx.upload(x, count * 8)); PG_TRY(dy.upload(y, count * 8)); PG_TRY(dr.zero(NX * NY * 8));
    int frac = std::max(8, std::min(60, 62 - ilog2(count + 1))); // same format rule as picgolf_create
    stage_cic_deposit_kernel<<<grid1(count), 256>>>(dx.as<double>(), dy.as<double>(), count, (int)NX, (int)NY, ldexp(1.0, frac),
                                                    dr.as<unsigned long long>());
    PG_TRY(finish());
    std::vector<long long> fx((size_t)(NX * NY));
    PG_TRY(dr.download(fx.data(), NX * NY * 8));
    for (int64_t n = 0; n < NX * NY; ++n) rho[n] = (double)fx[n] * ldexp(1.0, -frac) * w; // as solve2d_rows_fwd
    return 0;
}

PG_API int picgolf_stage_cic_gather(const double *Ex, const double *Ey, int64_t NX, int64_t NY, const double *x,
                                    const double *y, int64_t count, double *ex, double *ey)
{
    if (!Ex || !Ey || !x || !y || !ex || !ey || count < 0 || NX < 1 || NY < 1) return fail(PICGOLF_ERR_ARG, "bad argument");
    PG_TRY(stage_ready());
    DevBuf dEx, dEy, dx, dy, dex, dey;
    PG_TRY(dEx.upload(Ex, NX * NY * 8)); PG_TRY(dEy.upload(Ey, NX * NY * 8));
    PG_TRY(dx.upload(x, count * 8)); PG_TRY(dy.upload(y, count * 8));
    PG_TRY(dex.alloc(count * 8)); PG_TRY(dey.alloc(count * 8));
    stage_cic_gather_kernel<<<grid1(count), 256>>>(dEx.as<double>(), dEy.as<double>(), (int)NX, (int)NY, dx.as<double>(),
                                                   dy.as<double>(), count, dex.as<double>(), dey.as<double>());
    PG_TRY(finish());
    PG_TRY(dex.download(ex, count * 8));
    return dey.download(ey, count * 8);
}

PG_API int picgolf_stage_boris(double *vx, double *vy, double *vz, const double *Ex, const double *Ey, int64_t count,
                               double dt, double B0)
{
    if (!vx || !vy || !vz || !Ex || !Ey || count < 0) return fail(PICGOLF_ERR_ARG, "bad argument");
    PG_TRY(stage_ready());
    DevBuf a, b, c, e1, e2;
    PG_TRY(a.upload(vx, count * 8)); PG_TRY(b.upload(vy, count * 8)); PG_TRY(c.upload(vz, count * 8));
    PG_TRY(e1.upload(Ex, count * 8)); PG_TRY(e2.upload(Ey, count * 8));
    double t1 = B0 * dt / 2, tscale = 2 / (1 + (t1 * t1 + 0.0 + 0.0));
    stage_boris_kernel<<<grid1(count), 256>>>(a.as<double>(), b.as<double>(), c.as<double>(), e1.as<double>(), e2.as<double>(),
                                             count, dt, t1, tscale);
    PG_TRY(finish());
    PG_TRY(a.download(vx, count * 8)); PG_TRY(b.download(vy, count * 8));
    return c.download(vz, count * 8);
}

PG_API int picgolf_stage_fp64_peak(double *tflops)
{
    if (!tflops) return fail(PICGOLF_ERR_ARG, "NULL argument");
    PG_TRY(stage_ready());
    int dev = 0, sms = 0;
    PG_CUDA(cudaGetDevice(&dev));
    PG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DevBuf out;
    const int blocks = sms * 8, iters = 8192;
    PG_TRY(out.alloc((size_t)blocks * 256 * 8));
    cudaEvent_t e0, e1;
    PG_CUDA(cudaEventCreate(&e0)); PG_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        PG_CUDA(cudaEventRecord(e0));
        fp64_peak_kernel<<<blocks, 256>>>(out.as<double>(), iters, 0.999999, 1e-9);
        PG_CUDA(cudaEventRecord(e1));
        PG_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        PG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double tf = 2.0 * 8.0 * iters * (double)blocks * 256.0 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    PG_TRY(finish());
    *tflops = best;
    return 0;
}

PG_API int picgolf_stage_quiet_start(int64_t P, int64_t first, int64_t count, double *x, double *v)
{
    if (!x || !v || count < 0 || first < 0 || first + count > P) return fail(PICGOLF_ERR_ARG, "bad argument");
    PG_TRY(stage_ready());
    DevBuf dx, dv;
    PG_TRY(dx.alloc(count * 8)); PG_TRY(dv.alloc(count * 8));
    quiet_start_kernel<<<(unsigned)std::min<int64_t>(grid1(count), 4096), 256>>>(dx.as<double>(), dv.as<double>(), count, first, P);
    PG_TRY(finish());
    PG_TRY(dx.download(x, count * 8));
    return dv.download(v, count * 8);
}

// ---- PIC2D3V.jl electrostatic path + omega-k post-processing (include/picgolf_es.h) ----
#include "picgolf_es.inc"
