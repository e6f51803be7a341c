"""Host-side mirror of the electrostatic path of the reference's `PIC2D3V` module (src/PIC2D3V.jl) over the C ABI
of include/picgolf_es.h -- same names, argument meaning and defaults, so a run reads like the Julia it replaces:

    shape     = BSplineWeighting(2)                      # PIC2D3V.BSplineWeighting{@stat 2}()      src/2D3V.jl:111
    electrons = Species(P, vth, n0, shape, Lx=Lx, Ly=Ly, charge=-1, mass=1)                          :114
    ions      = Species(P, vth / sqrt(M), n0, shape, Lx=Lx, Ly=Ly, charge=1, mass=M)                 :116
    field     = ElectrostaticField(NX, NY, Lx, Ly, dt=dt, B0x=B0)                                    :86
    diags     = ElectrostaticDiagnostics(NX, NY, NT, ntskip, 2)                                      :87
    sim = Simulation([electrons, ions], field, diags)
    sim.loop(NT)            # for t in 0:NT-1; loop!(plasma, field, to, t, _); diagnose!(diagnostics, field, plasma, t, to); end

Everything is computed by libpicgolf.so on the GPU (ctypes only; no CPU path, nothing here imports oracle/).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import ESConfig, PicGolfError, _check, _f64, _i64, _out_ptr, _vp, load

SHAPE_NGP, SHAPE_AREA, SHAPE_BSPLINE0 = 0, 1, 10


@dataclass(frozen=True)
class NGPWeighting:  # struct NGPWeighting <: AbstractShape   PIC2D3V.jl:142
    code: int = SHAPE_NGP


@dataclass(frozen=True)
class AreaWeighting:  # PIC2D3V.jl:143
    code: int = SHAPE_AREA


class BSplineWeighting:  # BSplineWeighting{N}   PIC2D3V.jl:145,1122-1188
    def __init__(self, N: int):
        if not 0 <= int(N) <= 5:
            raise ValueError("BSplineWeighting{N}: bspline is defined for N = 0..5 (PIC2D3V.jl:1124-1163)")
        self.N = int(N)
        self.code = SHAPE_BSPLINE0 + self.N


def calculateweight(n0, P, Lx, Ly):  # PIC2D3V.jl:189
    return n0 * Lx * Ly / P


class Species:
    """Species(P, vth, density, shape; Lx, Ly, charge=1, mass=1)   PIC2D3V.jl:194-213.
    Without `xyv` the Halton start of the reference is generated on the device when the Simulation is built;
    `xyv` (5 x P, rows x, y, vx, vy, vz -- Species.xyv, or its transpose P x 5 C-contiguous) supplies the state instead."""

    def __init__(self, P, vth, density, shape, *, Lx, Ly, charge=1, mass=1, xyv=None):
        self.P, self.vth, self.density, self.shape = int(P), float(vth), float(density), shape
        self.Lx, self.Ly, self.charge, self.mass = float(Lx), float(Ly), float(charge), float(mass)
        self.weight = calculateweight(self.density, self.P, self.Lx, self.Ly)
        self.xyv = None if xyv is None else np.asarray(xyv, dtype=np.float64)


class ElectrostaticField:
    """ElectrostaticField(NX, NY=NX, Lx=1, Ly=1; dt, B0x=0, B0y=0, B0z=0)   PIC2D3V.jl:284-292.
    accumulate=True is update! as written (Exy += real(E) every step, PIC2D3V.jl:294-297)."""

    def __init__(self, NX, NY=None, Lx=1.0, Ly=1.0, *, dt, B0x=0.0, B0y=0.0, B0z=0.0, accumulate=True):
        self.NX, self.NY = int(NX), int(NX if NY is None else NY)
        self.Lx, self.Ly, self.dt = float(Lx), float(Ly), float(dt)
        self.B0 = (float(B0x), float(B0y), float(B0z))
        self.accumulate = bool(accumulate)


class ElectrostaticDiagnostics:
    """ElectrostaticDiagnostics(NX, NY, NT, ntskip, ngskip=1)   PIC2D3V.jl:95-102."""

    def __init__(self, NX, NY, NT, ntskip, ngskip=1, *, history=True):
        assert NT >= ntskip  # @assert NT >= ntskip
        assert ngskip >= 1 and (ngskip & (ngskip - 1)) == 0  # @assert ispow2(ngskip)
        self.NX, self.NY, self.NT, self.ntskip, self.ngskip, self.history = int(NX), int(NY), int(NT), int(ntskip), int(ngskip), bool(history)


class Simulation:
    """plasma + field + diagnostics on one GPU (or one shard of the particles per rank)."""

    def __init__(self, plasma: Sequence[Species], field: ElectrostaticField, diagnostics: ElectrostaticDiagnostics, *, device=-1,
                 rank=0, nranks=1, sort_every=0):
        self._lib = load()
        self._h = _vp()
        cfg = ESConfig()
        cfg.struct_size = C.sizeof(ESConfig)
        cfg.nspecies = len(plasma)
        cfg.NX, cfg.NY, cfg.Lx, cfg.Ly, cfg.dt = field.NX, field.NY, field.Lx, field.Ly, field.dt
        cfg.B0x, cfg.B0y, cfg.B0z = field.B0
        cfg.NT, cfg.ntskip, cfg.ngskip = diagnostics.NT, diagnostics.ntskip, diagnostics.ngskip
        cfg.field_accumulate, cfg.field_history = int(field.accumulate), int(diagnostics.history)
        cfg.device, cfg.rank, cfg.nranks, cfg.sort_every = device, rank, nranks, sort_every
        if len(plasma) > 4:
            raise PicGolfError(-1, "at most 4 species")
        for s, sp in enumerate(plasma):
            cfg.species_P[s], cfg.species_shape[s] = sp.P, sp.shape.code
            cfg.species_charge[s], cfg.species_mass[s], cfg.species_weight[s] = sp.charge, sp.mass, sp.weight
        self.cfg, self.plasma, self.field, self.diagnostics = cfg, list(plasma), field, diagnostics
        _check(self._lib.picgolf_es_create(C.byref(cfg), C.byref(self._h)))
        self.ranges = []
        for s in range(len(plasma)):
            f, c = _i64(), _i64()
            _check(self._lib.picgolf_es_local_range(self._h, s, C.byref(f), C.byref(c)))
            self.ranges.append((f.value, c.value))
        self.ND = max(1, diagnostics.NT // diagnostics.ntskip)
        self.NXd, self.NYd = field.NX // diagnostics.ngskip, field.NY // diagnostics.ngskip
        if nranks == 1:
            self.init_particles()

    def init_particles(self):
        """Upload `xyv` of every species that has one (this rank's shard) or run the reference's Halton start on the device.
        With several ranks call it after connect(): the Halton start all-reduces its mean and variance."""
        for s, sp in enumerate(self.plasma):
            first, count = self.ranges[s]
            if sp.xyv is None:
                _check(self._lib.picgolf_es_init_species(self._h, s, sp.vth))
            else:
                a = sp.xyv
                if a.shape == (5, sp.P):  # Julia's 5 x P column-major == P x 5 C-contiguous
                    a = a.T
                if a.shape != (sp.P, 5):
                    raise PicGolfError(-1, f"species {s}: xyv must be 5 x P")
                _check(self._lib.picgolf_es_set_species_xyv(self._h, s, _f64(a[first:first + count]), count))

    # -- lifetime
    def close(self):
        if self._h:
            self._lib.picgolf_es_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- particles
    def set_species(self, s, x, y, vx, vy, vz):
        arrs = [_f64(a) for a in (x, y, vx, vy, vz)]
        _check(self._lib.picgolf_es_set_species(self._h, s, *arrs, arrs[0].size))

    def species(self, s):
        """(x, y, vx, vy, vz) of this rank's shard of species s."""
        n = self.ranges[s][1]
        out = [np.empty(n) for _ in range(5)]
        _check(self._lib.picgolf_es_get_species(self._h, s, *[_out_ptr(a) for a in out], n))
        return tuple(out)

    def xyv(self, s):
        """Species.xyv of the local shard as a P x 5 C-contiguous array (Julia: 5 x P)."""
        n = self.ranges[s][1]
        out = np.empty((n, 5))
        _check(self._lib.picgolf_es_get_species_xyv(self._h, s, out, n))
        return out

    # -- loop! + diagnose!
    def loop(self, nsteps: int = 1):
        _check(self._lib.picgolf_es_step(self._h, int(nsteps)))

    def synchronize(self):
        _check(self._lib.picgolf_es_synchronize(self._h))

    @property
    def steps_done(self) -> int:
        v = _i64()
        _check(self._lib.picgolf_es_steps_done(self._h, C.byref(v)))
        return v.value

    @property
    def launches(self) -> int:
        v = _i64()
        _check(self._lib.picgolf_es_launch_count(self._h, C.byref(v)))
        return v.value

    def sort_stats(self):
        """(sorts so far, particles that left their tile's shared-memory window) of the tile-sorted mode."""
        a, b = _i64(), _i64()
        _check(self._lib.picgolf_es_sort_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- fields and diagnostics
    def fields(self):
        """dict(rho, Ex, Ey, Exy_x, Exy_y), each NX x NY (Fortran order, [i, j] as in Julia)."""
        n = self.field.NX * self.field.NY
        out = {k: np.empty(n) for k in ("rho", "Ex", "Ey", "Exy_x", "Exy_y")}
        _check(self._lib.picgolf_es_get_fields(self._h, *[_out_ptr(out[k]) for k in ("rho", "Ex", "Ey", "Exy_x", "Exy_y")]))
        return {k: v.reshape((self.field.NX, self.field.NY), order="F") for k, v in out.items()}

    def set_field(self, Exy_x, Exy_y):
        _check(self._lib.picgolf_es_set_field(self._h, _f64(np.asarray(Exy_x).ravel(order="F")), _f64(np.asarray(Exy_y).ravel(order="F"))))

    def scalars(self):
        """dict(kineticenergy[rows], fieldenergy[rows], particlemomentum[rows, 3], characteristicmomentum[rows, 3])."""
        ke, fe = np.zeros(self.ND), np.zeros(self.ND)
        pm, cm = np.zeros((self.ND, 3)), np.zeros((self.ND, 3))
        rows = _i64()
        _check(self._lib.picgolf_es_get_diagnostics(self._h, _out_ptr(ke), _out_ptr(fe), _out_ptr(pm), _out_ptr(cm), self.ND, C.byref(rows)))
        r = rows.value
        return dict(kineticenergy=ke[:r], fieldenergy=fe[:r], particlemomentum=pm[:r], characteristicmomentum=cm[:r])

    def history(self, which: str):
        """Exs / Eys / phis as (NX/ngskip, NY/ngskip, slices), Fortran order."""
        w = {"Ex": 0, "Exs": 0, "Ey": 1, "Eys": 1, "phi": 2, "phis": 2}[which]
        out = np.zeros(self.NXd * self.NYd * self.ND)
        n = _i64()
        _check(self._lib.picgolf_es_get_field_history(self._h, w, _out_ptr(out), self.ND, C.byref(n)))
        return out[: self.NXd * self.NYd * n.value].reshape((self.NXd, self.NYd, n.value), order="F")

    def spectrum(self, which: str, axis: int = 0, mode: int = 1):
        """|fft| omega-k map of a stored history, (n, ND): mode 1 = abs.(fft(F))[:, 1, :] (PIC2D3V.jl:1483), mode 0 =
        sum(i -> abs.(fft(F[:, i, :])), 1:size(F, 2)) (Electrostatic2D3V.jl:219); axis 1 swaps the roles of x and y."""
        w = {"Ex": 0, "Exs": 0, "Ey": 1, "Eys": 1, "phi": 2, "phis": 2}[which]
        n = self.NXd if axis == 0 else self.NYd
        out = np.zeros(n * self.ND)
        _check(self._lib.picgolf_es_spectrum(self._h, w, axis, mode, out))
        return out.reshape((n, self.ND), order="F")

    # -- multi-GPU
    def connect(self):
        """One process per GPU: ship the NCCL id over torch.distributed (default group) and join; rho is all-reduced
        every step.  Then call init_particles()."""
        import torch.distributed as dist

        from . import comm_unique_id
        from .distributed import broadcast_bytes

        if self.cfg.nranks == 1:
            return
        uid = comm_unique_id() if dist.get_rank() == 0 else None
        uid = broadcast_bytes(uid, 128, src=0)
        buf = C.create_string_buffer(bytes(uid), 128)
        _check(self._lib.picgolf_es_comm_init(self._h, buf, self.cfg.nranks, self.cfg.rank))


# ---- stage-level helpers ------------------------------------------------------------------------------
def shape_weights(shape, z, NZ_Lz):
    """depositindicesfractions for an array of positions: (j0[count] 1-based unwrapped, wt[count, 6])."""
    z = _f64(np.atleast_1d(z))
    j0, wt = np.zeros(z.size, dtype=np.int32), np.zeros((z.size, 6))
    _check(load().picgolf_es_stage_shape(int(shape.code if hasattr(shape, "code") else shape), z, z.size, float(NZ_Lz), j0, wt))
    return j0, wt


def boris(vx, vy, vz, Ex, Ey, B0, dt, q_m):
    a, b, c = (_f64(np.atleast_1d(v)).copy() for v in (vx, vy, vz))
    ex, ey = _f64(np.broadcast_to(Ex, a.shape)).copy(), _f64(np.broadcast_to(Ey, a.shape)).copy()
    _check(load().picgolf_es_stage_boris(a, b, c, ex, ey, a.size, float(B0[0]), float(B0[1]), float(B0[2]), float(dt), float(q_m)))
    return a, b, c


def wk_spectrum(F, axis: int = 0, mode: int = 1):
    """|fft| omega-k map of a history F[NA, NB, ND] held by the caller (e.g. Exs of Electrostatic2D3V.jl:171)."""
    F = np.asarray(F, dtype=np.float64)
    NA, NB, ND = F.shape
    flat = _f64(F.ravel(order="F"))
    n = NA if axis == 0 else NB
    out = np.zeros(n * ND)
    _check(load().picgolf_stage_wk_spectrum(flat, NA, NB, ND, axis, mode, out))
    return out.reshape((n, ND), order="F")
