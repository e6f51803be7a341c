/*
 * picgolf.h -- C ABI of libpicgolf.so: the per-timestep particle-in-cell loop of
 * jwscook/ParticleInCellCodeGolf.jl as hand-written CUDA for NVIDIA B200 (sm_100a).
 *
 * The reference has no FFI seam: its scripts are top-level Julia whose `for t` loop body
 * reads and writes globals.  This ABI *is* the seam north_star prescribes: a Julia driver keeps
 * the script's parameter/initialisation lines and replaces the loop body by `ccall`s into this
 * library (driver/picgolf.jl).  Each entry point below cites the reference
 * lines (relative to the reference repo root) whose work it takes over.
 *
 * Conventions
 *   - plain C, no torch / C++ types; all arrays are caller-owned host `double*` (Julia
 *     `Vector{Float64}`), contiguous, copied in/out during the call only;
 *   - grids are 0-based here, 1-based in Julia: rho[k] <-> r[k+1]; 2D grids are NX x NY
 *     column-major (x fastest) as in src/Electrostatic2D3V.jl:67-69;
 *   - diagnostics D is T x ncols column-major (D[t,c] at (c-1)*T + (t-1)), Julia layout;
 *   - every function returns 0 on success or a negative picgolf_status; nothing throws across
 *     the ABI; "not converged after max_sweeps" is NOT an error (the reference proceeds
 *     silently, src/GaussianFixedPoint.jl:7);
 *   - there is no CPU fallback: without a CUDA device every compute entry returns
 *     PICGOLF_ERR_CUDA.
 */
#ifndef PICGOLF_H
#define PICGOLF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PICGOLF_VERSION 100 /* 0.1.0 */

typedef enum picgolf_status {
    PICGOLF_OK = 0,
    PICGOLF_ERR_ARG = -1,         /* bad argument / unsupported configuration */
    PICGOLF_ERR_CUDA = -2,        /* CUDA runtime error (see picgolf_last_error) */
    PICGOLF_ERR_NCCL = -3,        /* NCCL error or NCCL not loadable */
    PICGOLF_ERR_STATE = -4,       /* call order (e.g. step before particles were set) */
    PICGOLF_ERR_UNSUPPORTED = -5  /* valid in the reference, not built here (e.g. an erf-shape scheme on a grid that is not 2^k) */
} picgolf_status;

/* Which reference script's loop body the handle runs. */
typedef enum picgolf_scheme {
    PICGOLF_NGP_LEAPFROG = 1,     /* src/NGPFourier.jl:4-7 (+ NGPFourierWithDiagnostics.jl:6-7) */
    PICGOLF_GAUSS_LEAPFROG = 2,   /* src/Gaussian.jl:8-12 */
    PICGOLF_GAUSS_FIXEDPOINT = 3, /* src/GaussianFixedPoint.jl:7-12, src/GaussianFixedPointQuiet.jl:8-15 */
    PICGOLF_CIC_BORIS_2D3V = 4,   /* src/Electrostatic2D3V.jl:120-176 */
    PICGOLF_GAUSS_SIMPSON13 = 5,  /* src/GaussianFixedPointQuietSimpson13.jl:8-18 (Simpson-1/3 quadrature of E, 3 solves/sweep) */
    PICGOLF_AREA_SIMPSON13 = 6,   /* src/AreaFixedPointQuietSimpson13.jl:7-17 (same schedule, 2-cell "area" shape d(y) of line 5) */
    PICGOLF_GAUSS_BORIS_1D2V = 7, /* src/NGP1D2V.jl:39-63 (1D2V magnetised, erf shape +-7, Boris about z; Bernstein modes) */
    PICGOLF_GAUSS_BORIS_1D2V2S = 8 /* src/NGP1D2V2S.jl:31-53: two species of P particles each, stored one after the other
                                    * (global indices [0,P): q = -1, q/m = -1; [P,2P): q = +1, q/m = 1/mass_ratio);
                                    * same entry points as the 1D2V scheme, arrays of 2P */
} picgolf_scheme;

/* Deposit accumulation mode. */
typedef enum picgolf_deposit_mode {
    PICGOLF_DEPOSIT_AUTO = 0,     /* library picks (sorted windows when P/N is large) */
    PICGOLF_DEPOSIT_ATOMIC = 1,   /* shared-memory privatised grid + fp64 atomics, any particle order */
    PICGOLF_DEPOSIT_SORTED = 2,   /* cell-sorted particles, register/window accumulation per warp */
    PICGOLF_DEPOSIT_POLY = 3      /* Gaussian fixed point only: (cell, sign v)-sorted particles, cell-polynomial gather and
                                     moment deposit (pg_kernels_poly.cuh); AUTO picks it for >= 1024 particles per cell */
} picgolf_deposit_mode;

/*
 * Parameter surface of the scripts (src/NGPFourier.jl:1, src/GaussianFixedPoint.jl:1-5,
 * src/GaussianFixedPointQuiet.jl:1-6, src/Electrostatic2D3V.jl:23-25).  Zero-initialise, set
 * struct_size = sizeof(picgolf_config), or start from picgolf_config_default().
 */
typedef struct picgolf_config {
    int32_t struct_size;
    int32_t scheme;          /* picgolf_scheme */
    int64_t N;               /* 1D: grid cells N, 16..8192: a power of two, or -- NGP_LEAPFROG only, NGPFourier.jl's fft takes
                              * any N -- any even number (direct transforms instead of the radix FFT).  2D: NX (power of two) */
    int64_t NY;              /* 2D only */
    int64_t P;               /* GLOBAL particle count (all ranks); 1D2V2S: per species (the handle holds 2P particles) */
    int64_t T;               /* capacity of the diagnostics trace in rows (steps recorded) */
    double dt;               /* time step */
    double W;                /* mean charge density (rho averages to W); 2D: n0 */
    double w;                /* deposit weight per particle: 1D W/P*N; Gaussian.jl w/dx; 2D n0/P/(dx*dy) */
    double rtol;             /* fixed point: l (GaussianFixedPoint.jl:4 1e-8; Quiet.jl:5 4eps()) */
    double atol;             /* fixed point: atol (Quiet.jl:8 atol=0) */
    double B0;               /* 2D3V: magnetic field along x (Electrostatic2D3V.jl:24,32) */
    int32_t half_width;      /* Gaussian stencil half width: 6 (GaussianFixedPoint.jl:5) or 7 (Quiet.jl:6) */
    int32_t max_sweeps;      /* fixed point: 10 (`for _ in 0:9`) */
    int32_t diag_every;      /* record a diagnostics row every diag_every steps (2D: NS=2, :25); 1D: 1 */
    int32_t deposit_mode;    /* picgolf_deposit_mode */
    int32_t deterministic;   /* 1: bit-reproducible charge for ANY particle order, run and GPU count: every deposit is
                              * quantised to 64-bit fixed point and summed with integer atomics (order-free kernel).
                              * 0: the faster cell-sorted path may be used; it is reproducible between sweeps of a step
                              * but its summation order changes from run to run at round-off level. */
    int32_t sort_every;      /* re-sort particles by cell every this many steps (0 = library default) */
    int32_t device;          /* CUDA device ordinal; -1 = current device */
    int32_t rank, nranks;    /* particle sharding: this handle owns global indices [first, first+count) */
    int32_t reserved_;
    int64_t local_first;     /* first global particle index owned (0-based); -1 = even split by rank */
    int64_t local_count;     /* particles owned; -1 = even split by rank */
    double mass_ratio;       /* 1D2V2S: M, mass of a species-2 particle in units of species 1 (NGP1D2V2S.jl:13 M=8) */
    int32_t field_history;   /* 2D3V: 1 = keep the snapshots Exs, Eys, phis [NX, NY, T] of every recorded step on the device
                              * (Electrostatic2D3V.jl:171-173); read them with picgolf_get_snapshots_2d */
    int32_t reserved2_;
} picgolf_config;

typedef struct picgolf_handle_s *picgolf_handle;

/* ---- library ---------------------------------------------------------------------------- */
int picgolf_version(void);
/* Thread-local text of the last error on this host thread ("" if none). */
const char *picgolf_last_error(void);
/* Number of visible CUDA devices (0 if none; never an error). */
int picgolf_device_count(void);

/* Fill cfg with the literal parameters of a reference script:
 * scheme NGP_LEAPFROG   -> NGPFourier.jl:1   (N=128,P=64N,dt=1/4N,T=1024,W=200,w=W/P*N)
 * scheme GAUSS_LEAPFROG -> Gaussian.jl:2     (NX=128,NP=64NX,dt=1/10NX,NT=1024,W=1600,w=W/NP/dx)
 * scheme GAUSS_FIXEDPOINT, quiet=0 -> GaussianFixedPoint.jl:1-5 (N=128,P=32N,dt=1/6N,T=1024,W=400,hw 6,l=1e-8)
 * scheme GAUSS_FIXEDPOINT, quiet=1 -> GaussianFixedPointQuiet.jl:1-6 (N=64,P=32N,T=2^13,W=32pi^2/3,hw 7,l=4eps)
 * scheme CIC_BORIS_2D3V -> Electrostatic2D3V.jl:23-25 (NX=NY=128,P=NX*NY*2^5,T=2^13,n0=4pi^2,...)
 * scheme GAUSS_SIMPSON13 -> GaussianFixedPointQuietSimpson13.jl:1-6 (same literals as the quiet fixed point)
 * scheme GAUSS_BORIS_1D2V -> NGP1D2V.jl:22-23 (N=512,P=15N,T=TO=2^14/16 rows,n0=4pi^2,vth,dt,B0,w=n0/P, diag_every=16)
 * scheme GAUSS_BORIS_1D2V2S -> NGP1D2V2S.jl:13-14 (N=256,P=8N per species,M=8,T=TO=2^16/32=2048 rows,diag_every=T/TO=32,
 *        vth=sqrt(n0)/N/8,dt=1/N/16vth,B0=sqrt(n0)/8,w=n0/2P) */
int picgolf_config_default(picgolf_config *cfg, int scheme, int quiet);

/* ---- lifetime --------------------------------------------------------------------------- */
/* Allocates all device state (replaces the array allocations of NGPFourier.jl:2,
 * GaussianFixedPoint.jl:2-3,5, Electrostatic2D3V.jl:43-49,67-72). */
int picgolf_create(const picgolf_config *cfg, picgolf_handle *out);
int picgolf_destroy(picgolf_handle h);
/* Local shard owned by this handle. */
int picgolf_local_range(picgolf_handle h, int64_t *first, int64_t *count);

/* ---- particle state --------------------------------------------------------------------- */
/* 1D1V: x, v of the local shard (length local_count).  Replaces `x=rand(P); v=...`
 * (NGPFourier.jl:2, GaussianFixedPoint.jl:2) -- Julia's RNG stream is not reproducible
 * elsewhere, so random starts are passed in.  Also resets E to zeros and the step counter. */
int picgolf_set_particles(picgolf_handle h, const double *x, const double *v, int64_t count);
/* 2D3V: x,y in (0,1], vx,vy,vz (Electrostatic2D3V.jl:45-55). */
int picgolf_set_particles_2d3v(picgolf_handle h, const double *x, const double *y, const double *vx,
                               const double *vy, const double *vz, int64_t count);
/* 1D2V (NGP1D2V.jl:30-32): x, vx, vy of the local shard. */
int picgolf_set_particles_1d2v(picgolf_handle h, const double *x, const double *vx, const double *vy, int64_t count);
int picgolf_get_particles_1d2v(picgolf_handle h, double *x, double *vx, double *vy, int64_t count);
/* Bit-reversal quiet start on device: x=(bitreverse.(0:P-1).+2.0^63)/2.0^64, v=+-1 by halves
 * (GaussianFixedPointQuiet.jl:2-3), generated from the GLOBAL index so shards agree. */
int picgolf_init_quiet(picgolf_handle h);
/* Seeded synthetic two-stream start on device (counter-based splitmix64, NOT Julia's rand):
 * x ~ U[0,1), v = -1 for global j <= P/2 else +1 (the NGPFourier.jl:2 pattern).  For 2D3V:
 * x,y ~ U(0,1], v Maxwellian with per-component std vth/sqrt(2) (Electrostatic2D3V.jl:45-55
 * without the sample-mean correction).  1D: vth > 0 warms the beams, v = +-1 + vth*N(0,1) (the velocity spread of the
 * saturated two-stream state); vth = 0 is the cold start of the scripts. */
int picgolf_init_synthetic(picgolf_handle h, uint64_t seed, double vth);
/* Copy the local shard back in the caller's original particle order. */
int picgolf_get_particles(picgolf_handle h, double *x, double *v, int64_t count);
int picgolf_get_particles_2d3v(picgolf_handle h, double *x, double *y, double *vx, double *vy, double *vz,
                               int64_t count);

/* ---- the loop body ---------------------------------------------------------------------- */
/* Run nsteps time steps on the device without host synchronisation inside a step:
 *   NGP_LEAPFROG     NGPFourier.jl:5-6   u(); deposit; solve; u(); kick
 *   GAUSS_LEAPFROG   Gaussian.jl:9-10
 *   GAUSS_FIXEDPOINT GaussianFixedPoint.jl:7-10 (<= max_sweeps sweeps, 2-norm isapprox test on device)
 *   CIC_BORIS_2D3V   Electrostatic2D3V.jl:121-157
 *   GAUSS_SIMPSON13  GaussianFixedPointQuietSimpson13.jl:8-17 (fields: rho = rho(x,x), E = E[end,:])
 * and append one diagnostics row per recorded step.  Asynchronous: returns after enqueueing;
 * any getter or picgolf_synchronize() waits.
 * A call ends on the scripts' end-of-step state (x, v as the reference holds them after `for t`'s body; rho, E of the last solve),
 * whatever nsteps is: calling picgolf_step(h, 1) K times gives the bits of picgolf_step(h, K).  The charge of the NEXT step's first
 * solve is already deposited when a call returns (leapfrog schemes: at the half-drifted positions u() will produce; fixed point:
 * at the first mid-points), so K steps cost K particle passes however the calls are split. */
int picgolf_step(picgolf_handle h, int64_t nsteps);
int picgolf_synchronize(picgolf_handle h);
/* Streaming form of  picgolf_set_particles(x_in, v_in) -> picgolf_step(1) -> picgolf_get_particles(x_out, v_out)  for a driver
 * that pushes a sequence of particle states through the device (ensembles, parameter scans, a host-side outer loop around
 * the loop body of NGPFourier.jl:5-6 / GaussianFixedPoint.jl:7-10).  Same result as the three calls, but ASYNCHRONOUS and
 * pipelined: the library keeps three device buffer sets and two copy streams, so the upload of call n+1, the step of call n
 * and the download of call n-1 overlap (PCIe is full duplex) -- use pinned host memory (cudaHostRegister / cudaMallocHost),
 * pageable buffers serialise the copies.  x_in/v_in must stay untouched and x_out/v_out are valid only after
 * picgolf_synchronize() (or any later getter on the handle).  The field of the step starts from E = 0 like after
 * picgolf_set_particles; the diagnostics trace keeps growing by one row per call.  x_out / v_out may be NULL.
 * 2D3V: arrays in the order x, y, vx, vy, vz (Electrostatic2D3V.jl:125-138); a row is recorded when diag_every == 1. */
int picgolf_step_streamed(picgolf_handle h, const double *x_in, const double *v_in, double *x_out, double *v_out, int64_t count);
int picgolf_step_streamed_2d3v(picgolf_handle h, const double *const in[5], double *const out[5], int64_t count);
/* Steps completed since the particles were last set. */
int picgolf_steps_done(picgolf_handle h, int64_t *steps);

/* ---- fields and diagnostics ------------------------------------------------------------- */
/* 1D: rho (the last deposited charge density r / n) and E, length N each; either may be NULL.
 * (GaussianFixedPoint.jl:6,8; NGPFourier.jl:5). */
int picgolf_get_fields(picgolf_handle h, double *rho, double *E);
int picgolf_set_field(picgolf_handle h, const double *E); /* restore E (checkpoint/resume) */
/* 2D: rho, real(Ex), real(Ey), NX*NY each, column-major (Electrostatic2D3V.jl:67-69,171-172). */
int picgolf_get_fields_2d(picgolf_handle h, double *rho, double *Ex, double *Ey);
/* Diagnostics trace, column-major with leading dimension ld >= rows recorded.
 *   1D schemes: 4 columns D[t,1:4] exactly as GaussianFixedPoint.jl:10-11 forms them:
 *       D1 = sum(E.^2)/N/2*(2/W), D2 = sum(v.^2)*W/P/2*(2/W), D3 = D1+D2, D4 = sum(v/P).
 *       (NGPFourierWithDiagnostics.jl:6-7 is the same up to its `E.^2N` parse, see DESIGN.md.)
 *   2D3V: 5 columns K[ti,1:5] (Electrostatic2D3V.jl:166-170).
 *   1D2V: 5 columns D[ti,1:5] (NGP1D2V.jl:59-61,64; one row per window of diag_every = T/TO steps).
 * sweeps (may be NULL) receives the fixed-point sweep count of each recorded step.
 * rows_out receives the number of rows recorded. */
int picgolf_get_diagnostics(picgolf_handle h, double *D, int64_t ld, int32_t *sweeps, int64_t *rows_out);
/* 1D2V: time-averaged field history Es[N, rows] (NGP1D2V.jl:57,64: Es[:,ti] .+= E; Es ./= T/TO), column-major with
 * leading dimension N; cols_out = windows started so far.  Window length = diag_every steps. */
int picgolf_get_field_history(picgolf_handle h, double *Es, int64_t max_cols, int64_t *cols_out);
/* 2D3V with field_history = 1: the snapshots of Electrostatic2D3V.jl:171-173, one NX x NY slice per recorded step (every
 * diag_every = NS steps): which = 0 Exs (real.(Ex)), 1 Eys, 2 phis (real.(pifft * phi): phi is the spectrum of the charge density
 * with phi[1,1] = 0, so this is rho - mean(rho)).  out is NX x NY x slices column-major, ready for the omega-k maps of lines 219-233
 * (picgolf_stage_wk_spectrum, picgolf_es.h). */
int picgolf_get_snapshots_2d(picgolf_handle h, int which, double *out, int64_t max_slices, int64_t *slices_out);
/* Raw per-step sums (column-major, ld rows): 1D: sum(E.^2), sum(v.^2), sum(v), sweeps;
 * 2D: sum(Ex^2+Ey^2), sum(vx^2+vy^2), sum(vx), sum(vy).  Lets the driver form any K it wants. */
int picgolf_get_raw_diagnostics(picgolf_handle h, double *raw, int64_t ld, int64_t *rows_out);

/* CUDA-event time per stage accumulated since the last reset, milliseconds, in the reference's
 * TimerOutputs sections (Electrostatic2D3V.jl:121-175): [0] particle loop (deposit/gather/push),
 * [1] reduction (allreduce), [2] field solve, [3] sort, [4] total in picgolf_step.
 * Enabling timing (enable=1) adds event records; it never synchronises inside a step. */
int picgolf_stage_timing(picgolf_handle h, int enable);
int picgolf_stage_times(picgolf_handle h, double ms[5], int reset);
/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int picgolf_launch_count(picgolf_handle h, int64_t *launches);
/* Cell-sorted mode bookkeeping: number of sorts so far and number of particle deposits that fell
 * outside their warp's shared-memory window (slow path) -- a stale-sort indicator.
 * In polynomial mode (PICGOLF_DEPOSIT_POLY) the second number counts mid-stream flushes of a lane's moment set. */
int picgolf_sort_stats(picgolf_handle h, int64_t *sorts, int64_t *slow_particles);
/* How many of those sorts were fused into the particle passes of a step (polynomial mode: the final pass of the step writes
 * its results straight to their slots in the order of the next step's mid-points, no separate pass over the particles). */
int picgolf_fused_sorts(picgolf_handle h, int64_t *n);
/* Which deposit path the handle runs (a picgolf_deposit_mode other than AUTO: what AUTO resolved to). */
int picgolf_deposit_path(picgolf_handle h, int *mode);
/* The cudaStream_t the handle enqueues on (for CUDA-event timing by the caller). */
int picgolf_get_stream(picgolf_handle h, void **stream);

/* ---- multi-GPU (one process per GPU; particles shard, rho is all-reduced) ----------------- */
/* NCCL is resolved at run time from the process (dlsym) or libnccl.so.2 (dlopen).
 * Rank 0 calls picgolf_comm_unique_id and ships the 128 bytes to the others (torch.distributed,
 * MPI, a file); then every rank calls picgolf_comm_init. */
int picgolf_comm_unique_id(void *id128);
int picgolf_comm_init(picgolf_handle h, const void *id128, int nranks, int rank);
/* Optional, after picgolf_comm_init: sum the 1D charge grids over NVLink peer memory inside the solve kernel instead of
 * one ncclAllReduce per sweep (pg_peer.cuh; one process per GPU on ONE node).  Every rank calls picgolf_peer_export
 * (64-byte cudaIpc handle out), the nranks handles are gathered in rank order (torch.distributed / MPI), then every
 * rank calls picgolf_peer_connect.  picgolf_peer_status: whether it is in use, and whether a wait ever timed out
 * (a rank died: the results after that are undefined). */
int picgolf_peer_export(picgolf_handle h, void *handle64);
int picgolf_peer_connect(picgolf_handle h, const void *handles, int nranks, int rank);
int picgolf_peer_status(picgolf_handle h, int *enabled, int *timed_out);

/* ---- stage-level entry points (parity tests call these through the same ABI) -------------- */
/* f(x)=Int(mod1(round(x*N),N)) (NGPFourier.jl:3); idx1 is 1-based like Julia. */
int picgolf_stage_ngp_index(const double *x, int64_t count, int64_t N, int32_t *idx1);
/* Julia float mod(x,1) (NGPFourier.jl:2, GaussianFixedPoint.jl:9). */
int picgolf_stage_mod1(const double *x, int64_t count, double *out);
/* d(c): (mod1(i,N), ff(i,c)) for i in (-hw:hw).+Int(round(c*N)) (GaussianFixedPoint.jl:4-5);
 * idx1, wt are count x (2hw+1) row-major. */
int picgolf_stage_gauss_stencil(const double *c, int64_t count, int64_t N, int hw, int32_t *idx1, double *wt);
/* NGP deposit n[f(j)]+=w (NGPFourier.jl:5). */
int picgolf_stage_ngp_deposit(const double *x, int64_t count, int64_t N, double w, double *rho);
/* rho(x,y): deposit at (x+y)/2 (GaussianFixedPoint.jl:6). mode: picgolf_deposit_mode. */
int picgolf_stage_gauss_deposit(const double *x, const double *y, int64_t count, int64_t N, int hw, double w,
                                int mode, double *rho);
/* sum(k->E[k[1]]*k[2], d(c)) (GaussianFixedPoint.jl:9). */
int picgolf_stage_gauss_gather(const double *E, int64_t N, int hw, const double *c, int64_t count, double *out);
/* E=real.(ifft((xi=fft(rho)./ik; xi[1]*=0; xi))) (NGPFourier.jl:3,5; GaussianFixedPoint.jl:3,8). */
int picgolf_stage_solve1d(const double *rho, int64_t N, double *E);
/* 2D field invert + solve (Electrostatic2D3V.jl:70-81,142-157): real(Ex), real(Ey). */
int picgolf_stage_solve2d(const double *rho, int64_t NX, int64_t NY, double *Ex, double *Ey);
/* CIC deposit / gather (Electrostatic2D3V.jl:84-109). */
int picgolf_stage_cic_deposit(const double *x, const double *y, int64_t count, int64_t NX, int64_t NY, double w,
                              double *rho);
int picgolf_stage_cic_gather(const double *Ex, const double *Ey, int64_t NX, int64_t NY, const double *x,
                             const double *y, int64_t count, double *ex, double *ey);
/* boris() (Electrostatic2D3V.jl:32-41) on arrays, in place. */
int picgolf_stage_boris(double *vx, double *vy, double *vz, const double *Ex, const double *Ey, int64_t count,
                        double dt, double B0);
/* Measured FP64 FMA peak of the device (TFLOP/s, 2 flops per lane-FMA): the physically binding roofline of the
 * erf-shape kernels; MEASURED_PEAKS.json carries no fp64 figure. */
int picgolf_stage_fp64_peak(double *tflops);
/* Quiet start for global indices [first, first+count) of P (GaussianFixedPointQuiet.jl:2-3). */
int picgolf_stage_quiet_start(int64_t P, int64_t first, int64_t count, double *x, double *v);

#ifdef __cplusplus
}
#endif
#endif /* PICGOLF_H */
