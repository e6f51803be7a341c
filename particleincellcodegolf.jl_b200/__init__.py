"""particleincellcodegolf.jl_b200 -- host-side mirror of libpicgolf.so (include/picgolf.h).

The product is the CUDA library; this module only binds its C ABI with ctypes and keeps the
reference scripts' parameter surface (N, P, dt, T, W, w, l, stencil half width) so that a run reads
like the Julia it replaces:

    src/NGPFourier.jl            -> ngp_fourier(...)
    src/Gaussian.jl              -> gaussian(...)
    src/GaussianFixedPoint.jl    -> gaussian_fixed_point(...)
    src/GaussianFixedPointQuiet.jl -> gaussian_fixed_point_quiet(...)
    src/Electrostatic2D3V.jl     -> electrostatic_2d3v(...)
    src/PIC2D3V.jl (electrostatic path: Species, shapes, ElectrostaticField, ElectrostaticDiagnostics, loop!, diagnose!)
                                 -> the pic2d3v submodule

There is no CPU path: importing works anywhere (so the ABI can be checked), but every compute call
needs a CUDA device and raises PicGolfError otherwise.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "lib", "libpicgolf.so")
HEADER_PATH = os.path.join(_ROOT, "include", "picgolf.h")
ES_HEADER_PATH = os.path.join(_ROOT, "include", "picgolf_es.h")
HEADER_PATHS = [HEADER_PATH, ES_HEADER_PATH]

NGP_LEAPFROG, GAUSS_LEAPFROG, GAUSS_FIXEDPOINT, CIC_BORIS_2D3V, GAUSS_SIMPSON13, AREA_SIMPSON13, GAUSS_BORIS_1D2V = 1, 2, 3, 4, 5, 6, 7
GAUSS_BORIS_1D2V2S = 8
DEPOSIT_AUTO, DEPOSIT_ATOMIC, DEPOSIT_SORTED, DEPOSIT_POLY = 0, 1, 2, 3

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


class PicGolfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"picgolf error {code}: {msg}")
        self.code = code


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/picgolf.cu for sm_100a into lib/libpicgolf.so (in-tree, travels with gpurun)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in sorted(os.listdir(src_dir))] + HEADER_PATHS
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(src_dir, "picgolf.cu"), "-ldl"]
    subprocess.check_call(cmd)
    return LIB_PATH


class Config(C.Structure):
    """ctypes image of `picgolf_config` (include/picgolf.h)."""
    _fields_ = [
        ("struct_size", C.c_int32), ("scheme", C.c_int32),
        ("N", C.c_int64), ("NY", C.c_int64), ("P", C.c_int64), ("T", C.c_int64),
        ("dt", C.c_double), ("W", C.c_double), ("w", C.c_double), ("rtol", C.c_double), ("atol", C.c_double),
        ("B0", C.c_double),
        ("half_width", C.c_int32), ("max_sweeps", C.c_int32), ("diag_every", C.c_int32), ("deposit_mode", C.c_int32),
        ("deterministic", C.c_int32), ("sort_every", C.c_int32), ("device", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32), ("reserved_", C.c_int32),
        ("local_first", C.c_int64), ("local_count", C.c_int64), ("mass_ratio", C.c_double),
        ("field_history", C.c_int32), ("reserved2_", C.c_int32),
    ]


class ESConfig(C.Structure):
    """ctypes image of `picgolf_es_config` (include/picgolf_es.h)."""
    _fields_ = [
        ("struct_size", C.c_int32), ("nspecies", C.c_int32), ("NX", C.c_int64), ("NY", C.c_int64),
        ("Lx", C.c_double), ("Ly", C.c_double), ("dt", C.c_double), ("B0x", C.c_double), ("B0y", C.c_double), ("B0z", C.c_double),
        ("NT", C.c_int64), ("ntskip", C.c_int32), ("ngskip", C.c_int32), ("field_accumulate", C.c_int32),
        ("field_history", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("sort_every", C.c_int32),
        ("species_P", C.c_int64 * 4), ("species_shape", C.c_int32 * 4), ("species_charge", C.c_double * 4),
        ("species_mass", C.c_double * 4), ("species_weight", C.c_double * 4),
    ]


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_vp, _i64, _d, _int = C.c_void_p, C.c_int64, C.c_double, C.c_int

# name -> (argtypes) ; every function returns int except picgolf_last_error.  This table is checked
# against include/picgolf.h by tests/test_abi.py.
_SIGNATURES = {
    "picgolf_version": [],
    "picgolf_device_count": [],
    "picgolf_config_default": [C.POINTER(Config), _int, _int],
    "picgolf_create": [C.POINTER(Config), C.POINTER(_vp)],
    "picgolf_destroy": [_vp],
    "picgolf_local_range": [_vp, C.POINTER(_i64), C.POINTER(_i64)],
    "picgolf_step_streamed": [_vp, _vp, _vp, _vp, _vp, _i64],
    "picgolf_step_streamed_2d3v": [_vp, C.POINTER(_vp), C.POINTER(_vp), _i64],
    "picgolf_set_particles": [_vp, _dp, _dp, _i64],
    "picgolf_set_particles_2d3v": [_vp, _dp, _dp, _dp, _dp, _dp, _i64],
    "picgolf_set_particles_1d2v": [_vp, _dp, _dp, _dp, _i64],
    "picgolf_get_particles_1d2v": [_vp, _vp, _vp, _vp, _i64],
    "picgolf_get_field_history": [_vp, _vp, _i64, C.POINTER(_i64)],
    "picgolf_get_snapshots_2d": [_vp, _int, _vp, _i64, C.POINTER(_i64)],
    "picgolf_init_quiet": [_vp],
    "picgolf_init_synthetic": [_vp, C.c_uint64, _d],
    "picgolf_get_particles": [_vp, _vp, _vp, _i64],
    "picgolf_get_particles_2d3v": [_vp, _vp, _vp, _vp, _vp, _vp, _i64],
    "picgolf_step": [_vp, _i64],
    "picgolf_synchronize": [_vp],
    "picgolf_steps_done": [_vp, C.POINTER(_i64)],
    "picgolf_get_fields": [_vp, _vp, _vp],
    "picgolf_set_field": [_vp, _dp],
    "picgolf_get_fields_2d": [_vp, _vp, _vp, _vp],
    "picgolf_get_diagnostics": [_vp, _vp, _i64, _vp, C.POINTER(_i64)],
    "picgolf_get_raw_diagnostics": [_vp, _vp, _i64, C.POINTER(_i64)],
    "picgolf_stage_timing": [_vp, _int],
    "picgolf_stage_times": [_vp, C.POINTER(_d * 5), _int],
    "picgolf_launch_count": [_vp, C.POINTER(_i64)],
    "picgolf_sort_stats": [_vp, C.POINTER(_i64), C.POINTER(_i64)],
    "picgolf_fused_sorts": [_vp, C.POINTER(_i64)],
    "picgolf_deposit_path": [_vp, C.POINTER(C.c_int)],
    "picgolf_get_stream": [_vp, C.POINTER(_vp)],
    "picgolf_comm_unique_id": [_vp],
    "picgolf_comm_init": [_vp, _vp, _int, _int],
    "picgolf_peer_export": [_vp, _vp],
    "picgolf_peer_connect": [_vp, _vp, _int, _int],
    "picgolf_peer_status": [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "picgolf_stage_ngp_index": [_dp, _i64, _i64, _ip],
    "picgolf_stage_mod1": [_dp, _i64, _dp],
    "picgolf_stage_gauss_stencil": [_dp, _i64, _i64, _int, _ip, _dp],
    "picgolf_stage_ngp_deposit": [_dp, _i64, _i64, _d, _dp],
    "picgolf_stage_gauss_deposit": [_dp, _dp, _i64, _i64, _int, _d, _int, _dp],
    "picgolf_stage_gauss_gather": [_dp, _i64, _int, _dp, _i64, _dp],
    "picgolf_stage_solve1d": [_dp, _i64, _dp],
    "picgolf_stage_solve2d": [_dp, _i64, _i64, _dp, _dp],
    "picgolf_stage_cic_deposit": [_dp, _dp, _i64, _i64, _i64, _d, _dp],
    "picgolf_stage_cic_gather": [_dp, _dp, _i64, _i64, _dp, _dp, _i64, _dp, _dp],
    "picgolf_stage_boris": [_dp, _dp, _dp, _dp, _dp, _i64, _d, _d],
    "picgolf_stage_fp64_peak": [C.POINTER(_d)],
    "picgolf_stage_quiet_start": [_i64, _i64, _i64, _dp, _dp],
    # include/picgolf_es.h (PIC2D3V.jl electrostatic path; host mirror in pic2d3v.py)
    "picgolf_es_config_default": [C.POINTER(ESConfig)],
    "picgolf_es_create": [C.POINTER(ESConfig), C.POINTER(_vp)],
    "picgolf_es_destroy": [_vp],
    "picgolf_es_local_range": [_vp, _int, C.POINTER(_i64), C.POINTER(_i64)],
    "picgolf_es_set_species": [_vp, _int, _dp, _dp, _dp, _dp, _dp, _i64],
    "picgolf_es_get_species": [_vp, _int, _vp, _vp, _vp, _vp, _vp, _i64],
    "picgolf_es_set_species_xyv": [_vp, _int, _dp, _i64],
    "picgolf_es_get_species_xyv": [_vp, _int, _dp, _i64],
    "picgolf_es_init_species": [_vp, _int, _d],
    "picgolf_es_step": [_vp, _i64],
    "picgolf_es_synchronize": [_vp],
    "picgolf_es_steps_done": [_vp, C.POINTER(_i64)],
    "picgolf_es_get_fields": [_vp, _vp, _vp, _vp, _vp, _vp],
    "picgolf_es_set_field": [_vp, _dp, _dp],
    "picgolf_es_get_diagnostics": [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64)],
    "picgolf_es_get_field_history": [_vp, _int, _vp, _i64, C.POINTER(_i64)],
    "picgolf_es_spectrum": [_vp, _int, _int, _int, _dp],
    "picgolf_es_launch_count": [_vp, C.POINTER(_i64)],
    "picgolf_es_sort_stats": [_vp, C.POINTER(_i64), C.POINTER(_i64)],
    "picgolf_es_get_stream": [_vp, C.POINTER(_vp)],
    "picgolf_es_comm_init": [_vp, _vp, _int, _int],
    "picgolf_es_stage_shape": [_int, _dp, _i64, _d, _ip, _dp],
    "picgolf_es_stage_boris": [_dp, _dp, _dp, _dp, _dp, _i64, _d, _d, _d, _d, _d],
    "picgolf_stage_wk_spectrum": [_dp, _i64, _i64, _i64, _int, _int, _dp],
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen libpicgolf.so.  Fails loudly if it was not built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PicGolfError(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(nvcc, sm_100a). libpicgolf has no CPU or PyTorch fallback.")
        # PICGOLF_LIB: an alternative build of the same sources (kernel tuning A/B runs, tools/build_variants.sh)
        L = C.CDLL(os.environ.get("PICGOLF_LIB") or LIB_PATH)
        for name, args in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        L.picgolf_last_error.restype = C.c_char_p
        L.picgolf_last_error.argtypes = []
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise PicGolfError(rc, load().picgolf_last_error().decode())


def device_count() -> int:
    return load().picgolf_device_count()


def default_config(scheme: int, quiet: bool = False) -> Config:
    cfg = Config()
    _check(load().picgolf_config_default(C.byref(cfg), scheme, 1 if quiet else 0))
    return cfg


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _out_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_vp)


def shard_range(P: int, rank: int, nranks: int):
    """Contiguous global index range [first, first+count) of a rank (remainder to the low ranks);
    the same rule picgolf_create applies when local_first/local_count are -1."""
    q, r = divmod(P, nranks)
    first = q * rank + min(rank, r)
    return first, q + (1 if rank < r else 0)


class PIC:
    """One simulation handle: the state `x, v, E, rho, D` of a reference script living on the GPU."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self._h = _vp()
        self._lib = load()
        _check(self._lib.picgolf_create(C.byref(cfg), C.byref(self._h)))
        f, c = _i64(), _i64()
        _check(self._lib.picgolf_local_range(self._h, C.byref(f), C.byref(c)))
        self.first, self.count = f.value, c.value
        self.is2d = cfg.scheme == CIC_BORIS_2D3V
        self.is1d2v = cfg.scheme in (GAUSS_BORIS_1D2V, GAUSS_BORIS_1D2V2S)
        self.ncell = cfg.N * (cfg.NY if self.is2d else 1)

    # -- lifetime
    def close(self):
        if self._h:
            self._lib.picgolf_destroy(self._h)
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
    def set_particles(self, x, v, y=None, vy=None, vz=None):
        if self.is2d:
            arrs = [_f64(a) for a in (x, y, v, vy, vz)]  # x, y, vx, vy, vz
            _check(self._lib.picgolf_set_particles_2d3v(self._h, *arrs, arrs[0].size))
        elif self.is1d2v:
            arrs = [_f64(a) for a in (x, v, vy)]  # x, vx, vy
            _check(self._lib.picgolf_set_particles_1d2v(self._h, *arrs, arrs[0].size))
        else:
            x, v = _f64(x), _f64(v)
            _check(self._lib.picgolf_set_particles(self._h, x, v, x.size))

    def step_streamed(self, inputs, outputs):
        """picgolf_step_streamed: upload `inputs` (x, v) or (x, y, vx, vy, vz), one step, download into `outputs`, all
        asynchronous and pipelined over three device buffer sets.  The arrays are contiguous float64 numpy arrays (pinned host
        memory for the copies to overlap) that must stay alive, and `outputs` are valid only after synchronize()."""
        n = self.count
        for a in list(inputs) + list(outputs):
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == n
        if self.is2d:
            pin = (_vp * 5)(*[a.ctypes.data for a in inputs])
            pout = (_vp * 5)(*[a.ctypes.data for a in outputs])
            _check(self._lib.picgolf_step_streamed_2d3v(self._h, pin, pout, n))
        else:
            _check(self._lib.picgolf_step_streamed(self._h, inputs[0].ctypes.data, inputs[1].ctypes.data,
                                                   outputs[0].ctypes.data, outputs[1].ctypes.data, n))

    def init_quiet(self):
        _check(self._lib.picgolf_init_quiet(self._h))

    def init_synthetic(self, seed: int = 0, vth: float = 0.0):
        _check(self._lib.picgolf_init_synthetic(self._h, C.c_uint64(seed), float(vth)))

    def particles(self):
        """(x, v) or (x, y, vx, vy, vz) of the local shard, in the caller's original order."""
        n = self.count
        if self.is2d:
            out = [np.empty(n) for _ in range(5)]
            _check(self._lib.picgolf_get_particles_2d3v(self._h, *[_out_ptr(a) for a in out], n))
            return tuple(out)
        if self.is1d2v:
            out = [np.empty(n) for _ in range(3)]
            _check(self._lib.picgolf_get_particles_1d2v(self._h, *[_out_ptr(a) for a in out], n))
            return tuple(out)
        x, v = np.empty(n), np.empty(n)
        _check(self._lib.picgolf_get_particles(self._h, _out_ptr(x), _out_ptr(v), n))
        return x, v

    # -- loop body
    def step(self, nsteps: int = 1):
        _check(self._lib.picgolf_step(self._h, int(nsteps)))

    def synchronize(self):
        _check(self._lib.picgolf_synchronize(self._h))

    @property
    def steps_done(self) -> int:
        s = _i64()
        _check(self._lib.picgolf_steps_done(self._h, C.byref(s)))
        return s.value

    # -- fields / diagnostics
    def fields(self):
        """1D: (rho, E).  2D: (rho, Ex, Ey) as NX x NY Fortran-ordered arrays."""
        if self.is2d:
            rho, ex, ey = (np.empty(self.ncell) for _ in range(3))
            _check(self._lib.picgolf_get_fields_2d(self._h, _out_ptr(rho), _out_ptr(ex), _out_ptr(ey)))
            shp = (self.cfg.N, self.cfg.NY)
            return tuple(a.reshape(shp, order="F") for a in (rho, ex, ey))
        rho, E = np.empty(self.ncell), np.empty(self.ncell)
        _check(self._lib.picgolf_get_fields(self._h, _out_ptr(rho), _out_ptr(E)))
        return rho, E

    def set_field(self, E):
        _check(self._lib.picgolf_set_field(self._h, _f64(E)))

    def diagnostics(self):
        """(D, sweeps): D is rows x 4 (1D, GaussianFixedPoint.jl:10-11) or rows x 5 (2D K)."""
        rows = _i64()
        _check(self._lib.picgolf_get_diagnostics(self._h, None, 0, None, C.byref(rows)))
        n, ncol = max(rows.value, 1), 5 if (self.is2d or self.is1d2v) else 4
        D = np.zeros((n, ncol), order="F")
        sw = np.zeros(n, dtype=np.int32)
        _check(self._lib.picgolf_get_diagnostics(self._h, _out_ptr(D), n, _out_ptr(sw), C.byref(rows)))
        return D[: rows.value], sw[: rows.value]

    def field_history(self):
        """1D2V: time-averaged field history Es (N x windows), NGP1D2V.jl:57,64."""
        cols = _i64()
        _check(self._lib.picgolf_get_field_history(self._h, None, 0, C.byref(cols)))
        Es = np.zeros((self.cfg.N, max(cols.value, 1)), order="F")
        _check(self._lib.picgolf_get_field_history(self._h, _out_ptr(Es), Es.shape[1], C.byref(cols)))
        return Es[:, : cols.value]

    def snapshots(self, which: str):
        """2D3V created with field_history=1: Exs / Eys / phis of src/Electrostatic2D3V.jl:171-173 as (NX, NY, slices), Fortran order."""
        w = {"Ex": 0, "Exs": 0, "Ey": 1, "Eys": 1, "phi": 2, "phis": 2}[which]
        n = _i64()
        _check(self._lib.picgolf_get_snapshots_2d(self._h, w, None, 0, C.byref(n)))
        out = np.zeros(self.ncell * max(n.value, 1))
        _check(self._lib.picgolf_get_snapshots_2d(self._h, w, _out_ptr(out), max(n.value, 1), C.byref(n)))
        return out[: self.ncell * n.value].reshape((self.cfg.N, self.cfg.NY, n.value), order="F")

    def raw_diagnostics(self):
        rows = _i64()
        _check(self._lib.picgolf_get_raw_diagnostics(self._h, None, 0, C.byref(rows)))
        n = max(rows.value, 1)
        R = np.zeros((n, 4), order="F")
        _check(self._lib.picgolf_get_raw_diagnostics(self._h, _out_ptr(R), n, C.byref(rows)))
        return R[: rows.value]

    # -- instrumentation
    def stage_timing(self, enable: bool = True):
        _check(self._lib.picgolf_stage_timing(self._h, 1 if enable else 0))

    def stage_times(self, reset: bool = False):
        ms = (_d * 5)()
        _check(self._lib.picgolf_stage_times(self._h, C.byref(ms), 1 if reset else 0))
        return dict(zip(("particles", "reduction", "solve", "sort", "total"), list(ms)))

    @property
    def launches(self) -> int:
        n = _i64()
        _check(self._lib.picgolf_launch_count(self._h, C.byref(n)))
        return n.value

    def sort_stats(self):
        """(number of cell sorts, particle deposits that took the out-of-window slow path -- polynomial mode: mid-stream
        flushes of a lane's moment set)."""
        a, b = _i64(), _i64()
        _check(self._lib.picgolf_sort_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def fused_sorts(self) -> int:
        """How many of the sorts were fused into the particle passes of a step (polynomial mode)."""
        n = _i64()
        _check(self._lib.picgolf_fused_sorts(self._h, C.byref(n)))
        return n.value

    @property
    def deposit_path(self) -> int:
        """DEPOSIT_ATOMIC / DEPOSIT_SORTED / DEPOSIT_POLY: what the handle actually runs (AUTO resolved)."""
        m = C.c_int()
        _check(self._lib.picgolf_deposit_path(self._h, C.byref(m)))
        return m.value

    @property
    def stream(self) -> int:
        s = _vp()
        _check(self._lib.picgolf_get_stream(self._h, C.byref(s)))
        return s.value or 0

    # -- multi-GPU
    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _check(self._lib.picgolf_comm_init(self._h, buf, self.cfg.nranks, self.cfg.rank))

    def peer_export(self) -> bytes:
        """64-byte cudaIpc handle of this rank's published charge grid (pg_peer.cuh)."""
        buf = C.create_string_buffer(64)
        _check(self._lib.picgolf_peer_export(self._h, buf))
        return buf.raw

    def peer_connect(self, handles: bytes) -> None:
        """handles: the nranks 64-byte handles in rank order."""
        assert len(handles) == 64 * self.cfg.nranks
        buf = C.create_string_buffer(handles, len(handles))
        _check(self._lib.picgolf_peer_connect(self._h, buf, self.cfg.nranks, self.cfg.rank))

    @property
    def peer_status(self):
        """(peer-memory reduction in use, a wait timed out)."""
        a, b = C.c_int(), C.c_int()
        _check(self._lib.picgolf_peer_status(self._h, C.byref(a), C.byref(b)))
        return bool(a.value), bool(b.value)


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(load().picgolf_comm_unique_id(buf))
    return buf.raw


# ------------------------------------------------------------------------------------------------
# Script-shaped constructors (same literals as line 1 of each reference script)
# ------------------------------------------------------------------------------------------------
def _finish(cfg: Config, rank, nranks, device, T, **over) -> PIC:
    cfg.rank, cfg.nranks, cfg.device = rank, nranks, device
    if T is not None:
        cfg.T = T
    for k, v in over.items():
        setattr(cfg, k, v)
    return PIC(cfg)


def ngp_fourier(N=128, P=None, dt=None, NT=1024, W=200.0, rank=0, nranks=1, device=-1, **over) -> PIC:
    """src/NGPFourier.jl:1  N=128;P=64N;dt=1/4N;NT=1024;W=200;w=W/P*N"""
    cfg = default_config(NGP_LEAPFROG)
    cfg.N = N
    cfg.P = 64 * N if P is None else P
    cfg.dt = 1 / (4 * N) if dt is None else dt
    cfg.W = W
    cfg.w = W / cfg.P * N
    return _finish(cfg, rank, nranks, device, NT, **over)


def gaussian(NX=128, NP=None, dt=None, NT=1024, W=1600.0, rank=0, nranks=1, device=-1, **over) -> PIC:
    """src/Gaussian.jl:2  NX=128; NP=64NX; dt=1/10NX; W=1600; w=W/NP; dx=1/NX (deposit scale w/dx)"""
    cfg = default_config(GAUSS_LEAPFROG)
    cfg.N = NX
    cfg.P = 64 * NX if NP is None else NP
    cfg.dt = 1 / (10 * NX) if dt is None else dt
    cfg.W = W
    cfg.w = W / cfg.P / (1 / NX)
    return _finish(cfg, rank, nranks, device, NT, **over)


def gaussian_fixed_point(N=128, P=None, dt=None, T=1024, W=400.0, l=1e-8, atol=0.0, half_width=6, max_sweeps=10,
                         rank=0, nranks=1, device=-1, **over) -> PIC:
    """src/GaussianFixedPoint.jl:1-5  N=128;P=32N;dt=1/6N;T=1024;W=400;w=W/P*N;l=1e-8; stencil -6:6"""
    cfg = default_config(GAUSS_FIXEDPOINT)
    cfg.N = N
    cfg.P = 32 * N if P is None else P
    cfg.dt = 1 / (6 * N) if dt is None else dt
    cfg.W = W
    cfg.w = W / cfg.P * N
    cfg.rtol, cfg.atol, cfg.half_width, cfg.max_sweeps = l, atol, half_width, max_sweeps
    return _finish(cfg, rank, nranks, device, T, **over)


def gaussian_fixed_point_quiet(N=64, P=None, dt=None, T=2 ** 13, W=32 * math.pi ** 2 / 3, l=4 * np.finfo(float).eps,
                               rank=0, nranks=1, device=-1, **over) -> PIC:
    """src/GaussianFixedPointQuiet.jl:1-6  N=64;P=32N;dt=1/6N;T=2^13;W=32pi^2/3;l=4eps();atol=0; stencil -7:7.
    Call .init_quiet() for the bit-reversal start (lines 2-3)."""
    return gaussian_fixed_point(N=N, P=P, dt=dt, T=T, W=W, l=l, atol=0.0, half_width=7, rank=rank, nranks=nranks,
                                device=device, **over)


def gaussian_fixed_point_quiet_simpson13(N=64, P=None, dt=None, T=2 ** 13, W=32 * math.pi ** 2 / 3, l=4 * np.finfo(float).eps,
                                         half_width=7, max_sweeps=10, rank=0, nranks=1, device=-1, **over) -> PIC:
    """src/GaussianFixedPointQuietSimpson13.jl:1-6 (same literals as the quiet fixed point); Simpson-1/3 time
    quadrature of E with three field solves per sweep (lines 8-18).  fields() returns (rho(x,x), E[end,:])."""
    cfg = default_config(GAUSS_SIMPSON13)
    cfg.N = N
    cfg.P = 32 * N if P is None else P
    cfg.dt = 1 / (6 * N) if dt is None else dt
    cfg.W = W
    cfg.w = W / cfg.P * N
    cfg.rtol, cfg.atol, cfg.half_width, cfg.max_sweeps = l, 0.0, half_width, max_sweeps
    return _finish(cfg, rank, nranks, device, T, **over)


def area_fixed_point_quiet_simpson13(N=64, P=None, dt=None, T=2 ** 13, W=32 * math.pi ** 2 / 3, l=1e-14, max_sweeps=10,
                                     rank=0, nranks=1, device=-1, **over) -> PIC:
    """src/AreaFixedPointQuietSimpson13.jl:1-5: the Simpson-1/3 schedule with the 2-cell area shape
    d(y)=(i=Int(mod1(ceil(y*N),N));o=ceil(y*N)-y*N;((i,1-o),(mod1(i-1,N),o))), l=1e-14."""
    cfg = default_config(AREA_SIMPSON13)
    cfg.N = N
    cfg.P = 32 * N if P is None else P
    cfg.dt = 1 / (6 * N) if dt is None else dt
    cfg.W = W
    cfg.w = W / cfg.P * N
    cfg.rtol, cfg.atol, cfg.max_sweeps = l, 0.0, max_sweeps
    return _finish(cfg, rank, nranks, device, T, **over)


def ngp_1d2v(N=512, P=None, T=2 ** 14, TO=None, n0=4 * math.pi ** 2, rank=0, nranks=1, device=-1, **over) -> PIC:
    """src/NGP1D2V.jl:22-23  N=512;P=15N;T=2^14;TO=T/16;n0=4pi^2;vth=sqrt(n0)/N/4;dt=1/N/6vth;B0=sqrt(n0)/16;w=n0/P.
    One diagnostics row / field-history column per window of T/TO steps.  set_particles(x, vx, vy=vy)."""
    cfg = default_config(GAUSS_BORIS_1D2V)
    TO = T // 16 if TO is None else TO
    cfg.N = N
    cfg.P = 15 * N if P is None else P
    vth = math.sqrt(n0) / N / 4
    cfg.W = n0
    cfg.dt = 1 / N / (6 * vth)
    cfg.B0 = math.sqrt(n0) / 16
    cfg.w = n0 / cfg.P
    cfg.diag_every = max(1, T // TO)
    cfg.half_width = 7
    pic = _finish(cfg, rank, nranks, device, TO, **over)
    pic.vth = vth
    return pic


def ngp_1d2v_2s(N=256, P=None, T=2 ** 16, TO=None, M=8.0, n0=4 * math.pi ** 2, rank=0, nranks=1, device=-1, **over) -> PIC:
    """src/NGP1D2V2S.jl:13-14  N=256;P=8N;T=2^16;TO=T/32;M=8;n0=4pi^2;vth=sqrt(n0)/N/8;dt=1/N/16vth;B0=sqrt(n0)/8;w=n0/2P.
    P particles per species; set_particles(x, vx, vy=vy) takes the two species one after the other ([x1; x2], ...)."""
    cfg = default_config(GAUSS_BORIS_1D2V2S)
    TO = T // 32 if TO is None else TO
    cfg.N = N
    cfg.P = 8 * N if P is None else P
    vth = math.sqrt(n0) / N / 8
    cfg.W = n0
    cfg.dt = 1 / N / (16 * vth)
    cfg.B0 = math.sqrt(n0) / 8
    cfg.w = n0 / (2 * cfg.P)
    cfg.mass_ratio = M
    cfg.diag_every = max(1, T // TO)
    cfg.half_width = 7
    pic = _finish(cfg, rank, nranks, device, TO, **over)
    pic.vth = vth
    return pic


def electrostatic_2d3v(NX=128, NY=None, P=None, T=2 ** 13, NS=2, n0=4 * math.pi ** 2, rank=0, nranks=1, device=-1,
                       **over) -> PIC:
    """src/Electrostatic2D3V.jl:23-25  NX=NY=128;P=NX*NY*2^5;NG=sqrt(NX^2+NY^2);n0=4pi^2;vth=sqrt(n0)/NG;
    dt=1/NG/6vth;B0=sqrt(n0)/4;NS=2;w=n0/P/(dx*dy).  T here is the number of diagnostics rows kept;
    field_history=1 also keeps the Exs/Eys/phis snapshots of lines 171-173 on the device (PIC.snapshots)."""
    cfg = default_config(CIC_BORIS_2D3V)
    NY = NX if NY is None else NY
    cfg.N, cfg.NY = NX, NY
    cfg.P = NX * NY * 32 if P is None else P
    NG = math.sqrt(NX ** 2 + NY ** 2)
    vth = math.sqrt(n0) / NG
    cfg.W = n0
    cfg.dt = 1 / NG / (6 * vth)
    cfg.B0 = math.sqrt(n0) / 4
    cfg.diag_every = NS
    cfg.w = n0 / cfg.P / ((1 / NX) * (1 / NY))
    pic = _finish(cfg, rank, nranks, device, T, **over)
    pic.vth = vth
    return pic


# ------------------------------------------------------------------------------------------------
# Stage-level calls (one reference expression each; see include/picgolf.h)
# ------------------------------------------------------------------------------------------------
def ngp_index(x, N: int) -> np.ndarray:
    x = _f64(np.atleast_1d(x))
    out = np.empty(x.size, dtype=np.int32)
    _check(load().picgolf_stage_ngp_index(x, x.size, N, out))
    return out


def mod1(x) -> np.ndarray:
    x = _f64(np.atleast_1d(x))
    out = np.empty_like(x)
    _check(load().picgolf_stage_mod1(x, x.size, out))
    return out


def gauss_stencil(c, N: int, hw: int = 6):
    c = _f64(np.atleast_1d(c))
    nw = 2 * hw + 1
    idx = np.empty((c.size, nw), dtype=np.int32)
    wt = np.empty((c.size, nw))
    _check(load().picgolf_stage_gauss_stencil(c, c.size, N, hw, idx.reshape(-1), wt.reshape(-1)))
    return idx, wt


def ngp_deposit(x, N: int, w: float) -> np.ndarray:
    x = _f64(x)
    rho = np.empty(N)
    _check(load().picgolf_stage_ngp_deposit(x, x.size, N, w, rho))
    return rho


def gauss_deposit(x, y, N: int, hw: int, w: float, mode: int = DEPOSIT_AUTO) -> np.ndarray:
    x, y = _f64(x), _f64(y)
    rho = np.empty(N)
    _check(load().picgolf_stage_gauss_deposit(x, y, x.size, N, hw, w, mode, rho))
    return rho


def gauss_gather(E, c, hw: int = 6) -> np.ndarray:
    E, c = _f64(E), _f64(np.atleast_1d(c))
    out = np.empty(c.size)
    _check(load().picgolf_stage_gauss_gather(E, E.size, hw, c, c.size, out))
    return out


def solve1d(rho) -> np.ndarray:
    rho = _f64(rho)
    E = np.empty_like(rho)
    _check(load().picgolf_stage_solve1d(rho, rho.size, E))
    return E


def solve2d(rho, NX: int, NY: int):
    rho = _f64(np.asarray(rho).reshape(-1, order="F"))
    Ex, Ey = np.empty(NX * NY), np.empty(NX * NY)
    _check(load().picgolf_stage_solve2d(rho, NX, NY, Ex, Ey))
    return Ex, Ey


def cic_deposit(x, y, NX: int, NY: int, w: float) -> np.ndarray:
    x, y = _f64(x), _f64(y)
    rho = np.empty(NX * NY)
    _check(load().picgolf_stage_cic_deposit(x, y, x.size, NX, NY, w, rho))
    return rho


def cic_gather(Ex, Ey, NX: int, NY: int, x, y):
    x, y = _f64(x), _f64(y)
    ex, ey = np.empty(x.size), np.empty(x.size)
    _check(load().picgolf_stage_cic_gather(_f64(Ex).reshape(-1), _f64(Ey).reshape(-1), NX, NY, x, y, x.size, ex, ey))
    return ex, ey


def boris(vx, vy, vz, Ex, Ey, dt: float, B0: float):
    vx, vy, vz = (_f64(a).copy() for a in (vx, vy, vz))
    _check(load().picgolf_stage_boris(vx, vy, vz, _f64(Ex), _f64(Ey), vx.size, dt, B0))
    return vx, vy, vz


def fp64_peak_tflops() -> float:
    """Measured FP64 FMA peak of the current device (TFLOP/s)."""
    t = _d()
    _check(load().picgolf_stage_fp64_peak(C.byref(t)))
    return t.value


def quiet_start(P: int, first: int = 0, count: Optional[int] = None):
    count = P - first if count is None else count
    x, v = np.empty(count), np.empty(count)
    _check(load().picgolf_stage_quiet_start(P, first, count, x, v))
    return x, v
