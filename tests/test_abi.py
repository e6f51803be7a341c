"""The C-ABI library loads, exports every symbol include/picgolf.h and include/picgolf_es.h declare, and refuses to compute
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest


def _declared(header):
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(picgolf_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(pg):
    lib = pg.load()
    names = sorted(set(n for hp in pg.HEADER_PATHS for n in _declared(hp)))  # include/picgolf.h + include/picgolf_es.h
    assert len(names) >= 58
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/*.h but not exported by libpicgolf.so"
    nm = subprocess.run(["nm", "-D", "--defined-only", pg.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (picgolf_\w+)", nm))
    assert set(names) <= exported
    # nothing but the C ABI is exported
    others = [l for l in nm.splitlines() if " T " in l and "picgolf_" not in l and "_init" not in l and "_fini" not in l]
    assert not others, others


def test_python_signature_table_matches_header(pg):
    names = set(n for hp in pg.HEADER_PATHS for n in _declared(hp))
    assert set(pg._SIGNATURES) | {"picgolf_last_error"} == names


def test_config_struct_layout(pg):
    cfg = pg.default_config(pg.GAUSS_FIXEDPOINT, quiet=False)
    assert cfg.struct_size == C.sizeof(pg.Config)  # C side wrote sizeof(picgolf_config)
    # src/GaussianFixedPoint.jl:1-5
    assert (cfg.N, cfg.P, cfg.T, cfg.half_width, cfg.max_sweeps) == (128, 4096, 1024, 6, 10)
    assert cfg.dt == 1 / (6 * 128) and cfg.W == 400 and cfg.w == 400 / 4096 * 128 and cfg.rtol == 1e-8
    q = pg.default_config(pg.GAUSS_FIXEDPOINT, quiet=True)  # src/GaussianFixedPointQuiet.jl:1-6
    assert (q.N, q.P, q.T, q.half_width) == (64, 2048, 8192, 7)
    assert q.rtol == 4 * np.finfo(float).eps and q.atol == 0
    assert abs(q.W - 32 * np.pi ** 2 / 3) < 1e-13
    n = pg.default_config(pg.NGP_LEAPFROG)  # src/NGPFourier.jl:1
    assert (n.N, n.P, n.T) == (128, 8192, 1024) and n.dt == 1 / 512 and n.w == 3.125
    g = pg.default_config(pg.GAUSS_LEAPFROG)  # src/Gaussian.jl:2
    assert g.dt == 1 / 1280 and g.w == 1600 / 8192 * 128
    e = pg.default_config(pg.CIC_BORIS_2D3V)  # src/Electrostatic2D3V.jl:23-25
    assert (e.N, e.NY, e.P, e.diag_every) == (128, 128, 128 * 128 * 32, 2)
    sp = pg.default_config(pg.GAUSS_SIMPSON13)  # src/GaussianFixedPointQuietSimpson13.jl:1-6
    assert (sp.N, sp.P, sp.T, sp.half_width, sp.scheme) == (64, 2048, 8192, 7, 5) and sp.rtol == 4 * np.finfo(float).eps
    b1 = pg.default_config(pg.GAUSS_BORIS_1D2V)  # src/NGP1D2V.jl:22: T=2^14, TO=T/16 rows, windows of T/TO = 16 steps
    assert (b1.N, b1.P, b1.T, b1.diag_every, b1.half_width) == (512, 15 * 512, 1024, 16, 7)
    b2 = pg.default_config(pg.GAUSS_BORIS_1D2V2S)  # src/NGP1D2V2S.jl:13: T=2^16, TO=T/32 = 2048 rows, windows of T/TO = 32 steps
    assert (b2.N, b2.P, b2.T, b2.diag_every, b2.mass_ratio) == (256, 8 * 256, 2048, 32, 8.0)


def test_es_config_struct_layout(pg):
    """picgolf_es_config (include/picgolf_es.h) and its ctypes image agree; defaults are the set-up of src/2D3V.jl:70-116."""
    cfg = pg.ESConfig()
    assert pg.load().picgolf_es_config_default(C.byref(cfg)) == 0
    assert cfg.struct_size == C.sizeof(pg.ESConfig) == 256
    assert (cfg.nspecies, cfg.NX, cfg.NY, cfg.NT, cfg.ntskip, cfg.ngskip) == (2, 128, 128, 1024, 4, 2)
    assert list(cfg.species_P)[:2] == [128 * 128 * 16] * 2 and list(cfg.species_shape)[:2] == [12, 12]
    assert list(cfg.species_charge)[:2] == [-1.0, 1.0] and list(cfg.species_mass)[:2] == [1.0, 32.0]
    n0 = 4 * np.pi ** 2
    assert cfg.species_weight[0] == n0 / (128 * 128 * 16) and abs(cfg.B0x - np.sqrt(n0) / 4) < 1e-15
    assert abs(cfg.dt - (1 / 128) / (6 * (1 / 128) * np.sqrt(n0))) < 1e-18 and cfg.field_accumulate == 1
    h = C.c_void_p()
    bad = pg.ESConfig.from_buffer_copy(cfg)
    bad.struct_size = 8
    assert pg.load().picgolf_es_create(C.byref(bad), C.byref(h)) == -1
    bad = pg.ESConfig.from_buffer_copy(cfg)
    bad.NX = 100
    assert pg.load().picgolf_es_create(C.byref(bad), C.byref(h)) == -5
    bad = pg.ESConfig.from_buffer_copy(cfg)
    bad.species_shape[1] = 16  # BSplineWeighting{6}: bspline is defined for 0..5 only
    assert pg.load().picgolf_es_create(C.byref(bad), C.byref(h)) == -1
    bad = pg.ESConfig.from_buffer_copy(cfg)
    bad.ngskip = 3  # @assert ispow2(ngskip)
    assert pg.load().picgolf_es_create(C.byref(bad), C.byref(h)) == -1
    assert pg.load().picgolf_es_step(None, 1) == -1 and pg.load().picgolf_es_destroy(None) == 0


def test_argument_errors_are_codes_not_crashes(pg):
    lib = pg.load()
    cfg = pg.default_config(pg.NGP_LEAPFROG)
    h = C.c_void_p()
    bad = pg.Config.from_buffer_copy(cfg)
    bad.struct_size = 8
    assert lib.picgolf_create(C.byref(bad), C.byref(h)) == -1
    assert b"struct_size" in lib.picgolf_last_error()
    bad = pg.Config.from_buffer_copy(cfg)
    bad.N = 101  # an odd grid: the reference's ik vector (NGPFourier.jl:3) is malformed there, not built here
    assert lib.picgolf_create(C.byref(bad), C.byref(h)) == -5
    bad.N = 100  # even, not a power of two: valid for the NGP leapfrog (direct transforms) -> stops only at the missing device
    assert lib.picgolf_create(C.byref(bad), C.byref(h)) == -2
    gfp = pg.default_config(pg.GAUSS_FIXEDPOINT)
    gfp.N = 100  # ... but not for the erf-shape schemes
    assert lib.picgolf_create(C.byref(gfp), C.byref(h)) == -5
    bad = pg.Config.from_buffer_copy(cfg)
    bad.scheme = 9
    assert lib.picgolf_create(C.byref(bad), C.byref(h)) == -1
    assert lib.picgolf_step(None, 1) == -1
    assert lib.picgolf_destroy(None) == 0


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback(pg):
    """Without a CUDA device every compute entry fails loudly with PICGOLF_ERR_CUDA."""
    assert pg.device_count() == 0
    with pytest.raises(pg.PicGolfError) as e:
        pg.PIC(pg.default_config(pg.NGP_LEAPFROG))
    assert e.value.code == -2 and "no CPU path" in str(e.value)
    with pytest.raises(pg.PicGolfError):
        pg.ngp_index([0.25], 128)
    with pytest.raises(pg.PicGolfError):
        pg.solve1d(np.ones(128))


def test_product_never_imports_the_oracle(pg):
    root = os.path.dirname(pg.HEADER_PATH)
    pkg = os.path.join(os.path.dirname(root), "particleincellcodegolf.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "picgolf_oracle" not in txt, f
