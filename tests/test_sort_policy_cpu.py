"""Re-sort interval policy of the polynomial passes (csrc/pg_sort_policy.h), compiled for the host: the decision is a pure function of the
probed flush rate and the age of the order, identical on every rank of a multi-GPU run.  Pins (1) that the expected flush rate of a sorted
stream does not depend on the number of ranks under weak scaling -- with the interval term not scaled by the rank count (as first written)
an 8-GPU run read its normal flush rate as 'hot' and ratcheted the interval down to a re-sort in every step --, (2) the cut / hold / grow
behaviour."""
import ctypes as C
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "particleincellcodegolf.jl_b200", "csrc")
SHIM = r'''
#include "pg_sort_policy.h"
extern "C" double expected(double N, int nsub, int nranks, double warps, double P, int det, int sublg) { return pg::poly_expected_flushes(N, nsub, nranks, warps, P, det != 0, sublg); }
extern "C" void probe(int *sort_every, int *grow_hold, int *quiet, double frac, int age, double expect) {
    pg::PolySortPolicy p{*sort_every, *grow_hold, *quiet != 0};
    pg::poly_sort_policy_probe(p, frac, age, expect);
    *sort_every = p.sort_every; *grow_hold = p.grow_hold; *quiet = p.quiet ? 1 : 0;
}
'''


@pytest.fixture(scope="module")
def lib():
    d = tempfile.mkdtemp()
    src, so = os.path.join(d, "shim.cpp"), os.path.join(d, "libpolicy.so")
    open(src, "w").write(SHIM)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", CSRC, "-o", so, src])
    L = C.CDLL(so)
    L.expected.restype = C.c_double
    L.expected.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
    L.probe.argtypes = [C.POINTER(C.c_int)] * 3 + [C.c_double, C.c_int, C.c_double]
    return L


class Policy:
    def __init__(self, lib, sort_every=16):
        self.lib, self.se, self.hold, self.quiet = lib, C.c_int(sort_every), C.c_int(0), C.c_int(0)

    def probe(self, frac, age, expect):
        self.lib.probe(C.byref(self.se), C.byref(self.hold), C.byref(self.quiet), frac, age, expect)
        return self.se.value


def test_expected_flush_rate_is_independent_of_the_rank_count(lib):
    N, per_gpu, warps_per_gpu = 4096, 1 << 28, 592 * 4
    e1 = lib.expected(N, 8, 1, warps_per_gpu, per_gpu, 0, 5)
    for r in (2, 4, 8):
        er = lib.expected(N, 8, r, warps_per_gpu * r, per_gpu * r, 0, 5)
        assert abs(er / e1 - 1) < 1e-12
    assert 0.02 < e1 < 0.04      # measured on B200 at 2^28 particles: 0.024 - 0.027 flushes per particle and step
    # a sorted cold stream must read as quiet, and far from 'hot', at any rank count
    for r in (1, 8):
        er = lib.expected(N, 8, r, warps_per_gpu * r, per_gpu * r, 0, 5)
        assert 0.0265 < 5e-5 + 1.5 * er and 0.0265 < 1e-3 + 3.0 * er


def test_cut_hold_and_growth(lib):
    expect = 0.03
    p = Policy(lib, 16)
    # a quiet step on an order that is not the oldest allowed changes nothing
    assert p.probe(0.02, 3, expect) == 16
    # quiet at the oldest allowed age: grow by half
    assert p.probe(0.02, 15, expect) == 24
    # a step that flushed a lot on a 5-step-old order: never that old again, and no growth for 32 steps
    assert p.probe(0.5, 5, expect) == 5
    for _ in range(31):
        assert p.probe(0.02, 4, expect) == 5
    assert p.probe(0.02, 4, expect) == 7          # the hold has run out
    # hot even on a fresh order: down to a re-sort in every step, and it stays there while steps stay hot
    assert p.probe(0.5, 0, expect) == 1
    assert p.probe(0.5, 0, expect) == 1
    # a hot probe of an order OLDER than the interval allows any more (the probe lags four steps) does not cut further
    p2 = Policy(lib, 4)
    assert p2.probe(0.5, 9, expect) == 4
    # the interval never exceeds 64
    p3 = Policy(lib, 60)
    assert p3.probe(0.0, 59, expect) == 64 and p3.probe(0.0, 63, expect) == 64
