// pg_sort_policy.h -- re-sort interval of the polynomial passes (host side, no CUDA): pure functions of the step number and of the flush
// counter, so that every rank of a multi-GPU run takes the same decision from the same (summed) counter.  Compiled into picgolf.cu
// and, for tests/test_sort_policy_cpu.py, into a small host library.
#pragma once
#include <algorithm>

namespace pg {

struct PolySortPolicy {
    int sort_every; // an order serves this many steps (fused re-sort: >= 1; stand-alone sort: >= 2)
    int grow_hold;  // steps left during which the interval may not grow (after a cut)
    bool quiet;     // the last probed step flushed no more than a sorted stream does
};

// Moment-set flushes per particle and step of a SORTED stream: in each of ~4 passes the warp that streams a (cell, sign v) group walks
// through its intervals and each of its 32 lanes hands in its set once per interval, plus once per warp range -- on EVERY rank, because
// every rank's shard covers the whole grid; P is the global particle count, warps the warps of all ranks.  Deterministic mode has no
// hysteresis: in the sub-bin that an interval edge cuts through (one in 2^sublg / nsub) a lane changes over between its two sets on up to
// every other particle, and the change-overs are counted too.
inline double poly_expected_flushes(double N, int nsub, int nranks, double warps_all_ranks, double P_global, bool det, int sublg)
{
    double expect = 4.0 * 32.0 * (2.0 * nsub * N * (double)nranks + warps_all_ranks) / P_global;
    if (det) expect += std::min(1.0, (double)nsub / (double)(1 << sublg));
    return expect;
}

// One probe of the fused-re-sort regime: `frac` flushes per particle were counted in a step that ran on an order `age` steps old.
// Never let the order get as old again as it was in a step that flushed a lot (down to a re-sort in every step); grow by half, up to 64,
// only when the oldest order the interval allows was still quiet, and not for 32 steps after a cut.
inline void poly_sort_policy_probe(PolySortPolicy &p, double frac, int age, double expect)
{
    p.quiet = frac < 5e-5 + 1.5 * expect;
    if (frac > 1e-3 + 3.0 * expect) {
        if (age < p.sort_every) { p.sort_every = std::max(1, age); p.grow_hold = 32; }
    } else if (p.quiet && age + 1 >= p.sort_every && p.grow_hold == 0) {
        p.sort_every = std::min(64, p.sort_every + std::max(1, p.sort_every / 2));
    }
    if (p.grow_hold > 0) --p.grow_hold;
}

} // namespace pg
