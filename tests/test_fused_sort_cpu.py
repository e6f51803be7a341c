"""Host-side restatement of the pieces of the re-sort that is fused into the polynomial passes (csrc/pg_kernels_poly.cuh:
cp_fs_key, cp_fs_runs, cp_fs_counts, cp_fs_scan_kernel): bit tricks and bookkeeping that can be pinned without a GPU.  The
device path itself is checked on the GPU (tests/test_parity_gpu.py::test_fused_resort_*)."""
import numpy as np
import pytest

MAGIC = 6755399441055744.0  # 1.5 * 2^52


def fs_key(xn, vb, N, sublg, dt):
    """cp_fs_key: bin of the next step's first mid-point xn + vb*dt/2 -- (cell centre, sign v, sub-cell position)."""
    scale = float(N << sublg)
    big = (vb * (dt / 2 * scale) + xn * scale) + (MAGIC + (1 << sublg) / 2)  # (fma on the device: same value up to one rounding)
    I = (big.view(np.int64) & 0xFFFFFFFF).astype(np.int64)
    I = np.where(I >= 1 << 31, I - (1 << 32), I)
    c = (I >> sublg) & (N - 1)
    return (((c << 1) | (vb >= 0)) << sublg) | (I & ((1 << sublg) - 1))


def test_key_is_the_bin_of_the_next_midpoint():
    N, sublg, dt = 64, 3, 1 / (6 * 64)
    rng = np.random.default_rng(1)
    x = rng.random(20000) * 1.2 - 0.1      # unwrapped end-of-step positions, a little outside [0, 1)
    v = rng.standard_normal(20000)
    key = fs_key(x, v, N, sublg, dt)
    assert key.min() >= 0 and key.max() < (2 * N) << sublg
    mid = x + v * dt / 2
    I = np.rint(mid * N * (1 << sublg)).astype(np.int64) + (1 << sublg) // 2   # position in sub-bins, offset by half a cell:
    cell, sub = I >> sublg, I & ((1 << sublg) - 1)                            # a cell's bins surround its stencil centre round(c*N)
    ok = ((key >> (sublg + 1)) == (cell & (N - 1))) & ((key & ((1 << sublg) - 1)) == sub)
    assert ok.mean() > 0.999                              # ties at bin edges only
    assert np.abs(cell - mid * N).max() <= 0.5 + 0.5 / (1 << sublg) + 1e-9
    assert np.array_equal((key >> sublg) & 1, (v >= 0).astype(np.int64))
    # monotone in the mid-point within a sign: what makes the slot order the order of the next step's deposits
    pos = v >= 0
    o = np.argsort(mid[pos])
    unwrapped = I[pos][o]
    assert np.all(np.diff(unwrapped) >= 0)


def fs_runs(keys):
    """cp_fs_runs for the 32 lanes of a warp, with the kernel's own bit operations."""
    head, length = [], []
    H = 0
    for lane in range(32):
        if lane == 0 or keys[lane] != keys[lane - 1]:
            H |= 1 << lane
    for lane in range(32):
        m = H & (0xFFFFFFFF >> (31 - lane))
        h = m.bit_length() - 1                            # 31 - clz
        above = H & ~((2 << h) - 1) & 0xFFFFFFFF
        nxt = (above & -above).bit_length() - 1 if above else 32   # ffs - 1
        head.append(h)
        length.append(nxt - h)
    return head, length


@pytest.mark.parametrize("seed", range(8))
def test_runs_partition_the_warp(seed):
    rng = np.random.default_rng(seed)
    keys = np.sort(rng.integers(0, 1 + seed, 32)) if seed % 2 else rng.integers(0, 3, 32)
    head, length = fs_runs(list(keys))
    covered = np.zeros(32, dtype=int)
    for lane in range(32):
        h, n = head[lane], length[lane]
        assert h <= lane < h + n and keys[h] == keys[lane]
        assert all(keys[q] == keys[h] for q in range(h, h + n))
        assert h == 0 or keys[h - 1] != keys[h]
        assert h + n == 32 or keys[h + n] != keys[h]
        if lane == h:
            covered[h:h + n] += 1
    assert np.all(covered == 1)                           # the heads' reservations cover every lane exactly once


def test_counting_rule_and_fallback():
    """cp_fs_counts(k): pass k counts iff k >= 1 and k + 1 >= fs_pred (the smaller sweep count of the last two steps).  A final
    sweep k >= 2 finds counts iff pass k-1 counted; otherwise the step writes unpermuted.  fs_pred = 0 (after a reset): always."""
    def counts(k, pred):
        return k >= 1 and k + 1 >= pred
    for pred in range(0, 11):
        for S in range(1, 11):                            # the step ends at sweep S
            scatter = S >= 2 and counts(S - 1, pred)
            if S >= 2 and (pred == 0 or S >= pred):
                assert scatter
            if S == 1 or (pred > 0 and S < pred):
                assert not scatter


def test_chunked_scan_equals_cumsum():
    """cp_fs_scan_kernel: 16 bins per thread, 1024 threads per block, block totals added up from the blocks below."""
    rng = np.random.default_rng(5)
    nbins = 1 << 18
    hist = rng.integers(0, 3000, nbins, dtype=np.int64)
    chunk = 1024 * 16
    totals = [hist[b:b + chunk].sum() for b in range(0, nbins, chunk)]
    cursor = np.empty(nbins, dtype=np.int64)
    for bi, b in enumerate(range(0, nbins, chunk)):
        below = sum(totals[:bi])
        c = hist[b:b + chunk].reshape(1024, 16)
        thread_excl = np.cumsum(c, axis=1) - c
        thread_tot = c.sum(axis=1)
        block_excl = np.cumsum(thread_tot) - thread_tot
        cursor[b:b + chunk] = (below + block_excl[:, None] + thread_excl).ravel()
    assert np.array_equal(cursor, np.cumsum(hist) - hist)
