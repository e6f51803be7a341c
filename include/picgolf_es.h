/*
 * picgolf_es.h -- C ABI of libpicgolf.so for the electrostatic path of the reference's un-golfed 2D3V module
 * src/PIC2D3V.jl and for the omega-k post-processing of the field histories (rank 4):
 *
 *   Species(P, vth, density, shape; Lx, Ly, charge, mass)        src/PIC2D3V.jl:148-213
 *   NGPWeighting / AreaWeighting / BSplineWeighting{0..5}          src/PIC2D3V.jl:141-145,1105-1188
 *   ElectrostaticField(NX, NY, Lx, Ly; dt, B0x, B0y, B0z)          src/PIC2D3V.jl:270-297
 *   ElectrostaticDiagnostics(NX, NY, NT, ntskip, ngskip)           src/PIC2D3V.jl:74-102
 *   loop!(plasma, field::ElectrostaticField, ...) ; diagnose!(...) src/PIC2D3V.jl:530-581,1301-1330
 *   fft of Exs/Eys/phis -> omega-k maps                            src/PIC2D3V.jl:1483-1493, src/Electrostatic2D3V.jl:219-233
 *
 * Same conventions as picgolf.h (plain C, caller-owned host double arrays copied in/out during the call, 0 or a negative
 * picgolf_status, picgolf_last_error() for the text, no CPU path).  A Julia driver keeps its Species / field
 * construction lines and replaces `loop!` + `diagnose!` inside `for t in 0:NT-1` (src/2D3V.jl:123-126) by
 * picgolf_es_step (driver/picgolf.jl).
 *
 * Layout: grids are NX x NY column-major (x fastest), 0-based here / 1-based in Julia; the halo ("offset") arrays of the
 * reference (rhos, Exy with buffer 3) are a CPU detail and are not exposed: the library keeps their periodic image and
 * wraps stencil indices with the reference's own unimod, which is the same arithmetic term by term.
 * Several species are stored one after the other; particles of a species shard over ranks by contiguous global index.
 */
#ifndef PICGOLF_ES_H
#define PICGOLF_ES_H

#include "picgolf.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PICGOLF_ES_MAX_SPECIES 4

/* AbstractShape of a species (src/PIC2D3V.jl:141-145). */
typedef enum picgolf_es_shape {
    PICGOLF_ES_NGP = 0,        /* NGPWeighting: ((i, 1),)                              :1115 */
    PICGOLF_ES_AREA = 1,       /* AreaWeighting: ((i, 1-r), (i+1, r))                  :1117-1119 */
    PICGOLF_ES_BSPLINE0 = 10,  /* BSplineWeighting{N} = PICGOLF_ES_BSPLINE0 + N, N = 0..5   :1122-1188 */
    PICGOLF_ES_BSPLINE1 = 11,
    PICGOLF_ES_BSPLINE2 = 12,
    PICGOLF_ES_BSPLINE3 = 13,
    PICGOLF_ES_BSPLINE4 = 14,
    PICGOLF_ES_BSPLINE5 = 15
} picgolf_es_shape;

typedef struct picgolf_es_config {
    int32_t struct_size;       /* sizeof(picgolf_es_config) */
    int32_t nspecies;          /* length of `plasma`, 1..PICGOLF_ES_MAX_SPECIES */
    int64_t NX, NY;            /* grid (powers of two, 16..1024) */
    double Lx, Ly;             /* box lengths (GridParameters :221-239) */
    double dt;                 /* time step (ElectrostaticField(...; dt)) */
    double B0x, B0y, B0z;      /* uniform magnetic field (ElectrostaticBoris :44-48) */
    int64_t NT;                /* steps the diagnostics are sized for: NT / ntskip rows (:95-99) */
    int32_t ntskip;            /* scalars every ntskip steps, field histories averaged over windows of ntskip steps */
    int32_t ngskip;            /* field histories keep every ngskip-th cell per dimension (power of two) */
    int32_t field_accumulate;  /* 1: update! as written -- Exy += real(E) every step and nothing resets it (:294-297, 21-27),
                                *    so the particles feel the running SUM of the solved fields;
                                * 0: Exy = real(E) of this step (what the module's other field types do, :475) */
    int32_t field_history;     /* 1: keep Exs, Eys, phis [NX/ngskip, NY/ngskip, NT/ntskip] on the device (:1322-1327) */
    int32_t device;            /* CUDA device ordinal; -1 = current */
    int32_t rank, nranks;      /* particle sharding (every species is split evenly over the ranks) */
    int32_t sort_every;        /* particle order: 0 = the library decides (shards of >= 2^20 particles with >= 8 per cell are kept
                                * sorted by 16x16-cell tile and run through shared-memory windows; re-sorted every 16 steps to begin with, then
                                * every 2...64 steps following the counted window misses -- fixed 16 on several GPUs);
                                * n > 0 = tile-sorted, re-sorted every n steps; < 0 = never sort (any-order kernel, global atomics).
                                * The reference sorts for cache locality too (sort!(s::Species, dx, dy) :210-215); getters always
                                * return the caller's particle order. */
    int64_t species_P[PICGOLF_ES_MAX_SPECIES];       /* GLOBAL particle count of each species */
    int32_t species_shape[PICGOLF_ES_MAX_SPECIES];   /* picgolf_es_shape */
    double species_charge[PICGOLF_ES_MAX_SPECIES];   /* Species.charge */
    double species_mass[PICGOLF_ES_MAX_SPECIES];     /* Species.mass */
    double species_weight[PICGOLF_ES_MAX_SPECIES];   /* Species.weight = calculateweight(n0, P, Lx, Ly) = n0*Lx*Ly/P  :189 */
} picgolf_es_config;

typedef struct picgolf_es_handle_s *picgolf_es_handle;

/* The commented-out electrostatic set-up of src/2D3V.jl:70-116: NX = NY = 128, Lx = Ly = 1, P = NX*NY*16 per species,
 * n0 = 4pi^2, dl = 1/NX, vth = dl*sqrt(n0), B0x = sqrt(n0)/4, dt = dl/6vth, ntskip = 4, ngskip = 2, NT = 2^10, electrons
 * (charge -1, mass 1) and ions (charge +1, mass 32), BSplineWeighting{2}, weight = n0*Lx*Ly/P, update! as written. */
int picgolf_es_config_default(picgolf_es_config *cfg);

/* ElectrostaticField(...) + ElectrostaticDiagnostics(...) + the particle storage of the species. */
int picgolf_es_create(const picgolf_es_config *cfg, picgolf_es_handle *out);
int picgolf_es_destroy(picgolf_es_handle h);
/* Global index range of species s owned by this handle. */
int picgolf_es_local_range(picgolf_es_handle h, int species, int64_t *first, int64_t *count);

/* Particle state of species s (local shard): five arrays (positions in (0, L], as Species(...) makes them) ... */
int picgolf_es_set_species(picgolf_es_handle h, int species, const double *x, const double *y, const double *vx,
                           const double *vy, const double *vz, int64_t count);
int picgolf_es_get_species(picgolf_es_handle h, int species, double *x, double *y, double *vx, double *vy, double *vz,
                           int64_t count);
/* ... or the reference's own storage, Species.xyv: a 5 x count column-major matrix (rows x, y, vx, vy, vz; :153,159-164). */
int picgolf_es_set_species_xyv(picgolf_es_handle h, int species, const double *xyv, int64_t count);
int picgolf_es_get_species_xyv(picgolf_es_handle h, int species, double *xyv, int64_t count);
/* Species(P, vth, density, shape; Lx, Ly) on the device (:194-213): Halton starts sample(P, b) = halton.(0:P-1, b, 1/sqrt(2))
 * in bases 2, 3 (positions) and 5, 7, 9 (velocities), v = vth*erfinv(2 sample - 1)*vth, mean removed, rescaled to the
 * corrected standard deviation vth/sqrt(2); generated from the global index, sums all-reduced (collective when nranks > 1). */
int picgolf_es_init_species(picgolf_es_handle h, int species, double vth);

/* nsteps x { loop!(plasma, field, to, t, _); diagnose!(diagnostics, field, plasma, t, to) }  (src/2D3V.jl:123-126),
 * t = steps done so far (starts at 0).  No host synchronisation inside; asynchronous like picgolf_step. */
int picgolf_es_step(picgolf_es_handle h, int64_t nsteps);
int picgolf_es_synchronize(picgolf_es_handle h);
int picgolf_es_steps_done(picgolf_es_handle h, int64_t *steps);

/* NX*NY each, any may be NULL: rho = the reduced charge density of the last step (phi before the transform, :557),
 * Ex, Ey = real(field.Ex), real(field.Ey) of the last solve, Exy_x, Exy_y = field.Exy[1|2, 1:NX, 1:NY]. */
int picgolf_es_get_fields(picgolf_es_handle h, double *rho, double *Ex, double *Ey, double *Exy_x, double *Exy_y);
/* Restore field.Exy (checkpoint/resume). */
int picgolf_es_set_field(picgolf_es_handle h, const double *Exy_x, const double *Exy_y);

/* ElectrostaticDiagnostics scalars, one entry per recorded row ti (t = (ti-1)*ntskip): kineticenergy[rows],
 * fieldenergy[rows] (= mean(abs2, Exy)/2 over the halo array, :1320), particlemomentum and characteristicmomentum
 * [3 x rows] column-major (component fastest).  Any output may be NULL.  Collective when nranks > 1. */
int picgolf_es_get_diagnostics(picgolf_es_handle h, double *kinetic, double *field, double *pmom, double *cmom,
                               int64_t max_rows, int64_t *rows_out);
/* Field histories (:1322-1327): which = 0 Exs, 1 Eys, 2 phis; out is (NX/ngskip) x (NY/ngskip) x slices column-major;
 * slice ti holds the average over steps (ti-1)*ntskip .. ti*ntskip-1 (a slice still being filled holds a partial sum).
 * "phis" is what the reference stores under that name: real(ifft(phi)) with phi the spectrum of rho, [1,1] zeroed. */
int picgolf_es_get_field_history(picgolf_es_handle h, int which, double *out, int64_t max_slices, int64_t *slices_out);
/* omega-k map of a stored history on the device (which as above; needs NT/ntskip a power of two >= 4).
 * axis 0: wavenumber along x, 1: along y.  mode 0: sum over the lines of the other axis of |fft2(F[line])|
 * (Electrostatic2D3V.jl:219,229); mode 1: |fft3(F)| on the k_other = 0 slice (PIC2D3V.jl:1483,1489).
 * out is n x slices column-major (n = NX/ngskip or NY/ngskip; wavenumber fastest), magnitudes before log10/slicing. */
int picgolf_es_spectrum(picgolf_es_handle h, int which, int axis, int mode, double *out);

int picgolf_es_launch_count(picgolf_es_handle h, int64_t *launches);
/* Tile-sorted mode: sorts so far, and particle deposits/gathers that fell outside their tile's shared-memory window
 * (slow path through global memory: a stale-sort indicator). */
int picgolf_es_sort_stats(picgolf_es_handle h, int64_t *sorts, int64_t *slow_particles);
int picgolf_es_get_stream(picgolf_es_handle h, void **stream);
/* One process per GPU: id128 from picgolf_comm_unique_id (picgolf.h); rho is all-reduced once per step. */
int picgolf_es_comm_init(picgolf_es_handle h, const void *id128, int nranks, int rank);

/* ---- stage-level entry points ------------------------------------------------------------- */
/* depositindicesfractions(shape, z, NZ, NZ_Lz) (:1105-1113): j0[p] = first grid index (1-based, unwrapped),
 * wt[6*p + k] = fraction k (zero beyond the shape's support). */
int picgolf_es_stage_shape(int shape, const double *z, int64_t count, double NZ_Lz, int32_t *j0, double *wt);
/* (boris::ElectrostaticBoris)(vx, vy, vz, Ex, Ey, q_m) on arrays, in place (:49-54). */
int picgolf_es_stage_boris(double *vx, double *vy, double *vz, const double *Ex, const double *Ey, int64_t count,
                           double B0x, double B0y, double B0z, double dt, double q_m);
/* The omega-k map of a caller-supplied history F[NA, NB, ND] (e.g. Exs of src/Electrostatic2D3V.jl:171 collected by the
 * driver): same axis/mode/out as picgolf_es_spectrum; NA, NB, ND powers of two, 4..8192. */
int picgolf_stage_wk_spectrum(const double *F, int64_t NA, int64_t NB, int64_t ND, int axis, int mode, double *out);

#ifdef __cplusplus
}
#endif
#endif /* PICGOLF_ES_H */
