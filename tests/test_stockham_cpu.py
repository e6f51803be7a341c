"""Index algebra of the register-blocked Stockham transform of the 1D solve (csrc/pg_fft.cuh: st_pass / fft_stockham8 /
solve1d_stock_kernel), restated in numpy: pass structure (radix-8 passes, a closing radix-4 / radix-2 pass), the
butterfly -> output map j = (i - k) R + k + r p with twiddles exp(-2 pi i k r / (p R)), the padded shared-memory layout,
and the inverse taken as the forward transform of the conjugate.  The device kernel is checked against the oracle on the
GPU (tests/test_parity_gpu.py::test_solve1d_parity); this pins the algorithm it implements."""
import numpy as np
import pytest


def st_phys(i):
    return i + (i >> 3)


def stockham(x):
    """Forward DFT by the passes of fft_stockham8: every 'thread' t < N/8 owns 8 points per pass."""
    N = x.size
    lg = N.bit_length() - 1
    buf = np.zeros(N + N // 8, dtype=complex)
    idx = st_phys(np.arange(N))
    assert np.unique(idx).size == N and idx.max() < buf.size  # the padding is a bijection into the buffer
    buf[idx] = x
    radices = [8] * (lg // 3) + ([4] if lg % 3 == 2 else [2] if lg % 3 == 1 else [])
    p = 1
    for R in radices:
        T = N // R
        i = np.arange(T)
        k = i & (p - 1)
        jb = (i - k) * R + k
        u = np.stack([buf[st_phys(i + r * T)] for r in range(R)])                 # [R, T] reads: contiguous in i
        w = np.exp(-2j * np.pi * np.outer(np.arange(R), k) / (p * R))             # twiddle of leg r: w^r, w = exp(-2 pi i k/(pR))
        u = u * w
        out = np.fft.fft(u, axis=0)                                               # the in-register DFT of R points, natural order
        new = buf.copy()
        for r in range(R):
            new[st_phys(jb + r * p)] = out[r]
        written = np.concatenate([st_phys(jb + r * p) for r in range(R)])
        assert np.unique(written).size == N                                       # every point written exactly once
        buf = new
        p *= R
    return buf[idx]


@pytest.mark.parametrize("N", [512, 1024, 2048, 4096, 8192])
def test_stockham_passes_are_the_dft(N):
    rng = np.random.default_rng(N)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    assert np.abs(stockham(x) - np.fft.fft(x)).max() < 1e-10 * np.abs(x).max() * N


@pytest.mark.parametrize("N", [512, 4096])
def test_solve_through_the_conjugate(N):
    """E = real(ifft(fft(rho)./ik, xi[1] = 0)) (src/NGPFourier.jl:3-5) == real(fft(conj(xi)))/N, with xi stored conjugated
    as solve1d_stock_kernel does: conj(z/(i b)) = (Im z)/b + i (Re z)/b."""
    rng = np.random.default_rng(3)
    rho = rng.standard_normal(N)
    ik = 2j * np.pi * np.concatenate(([1.0], np.arange(1, N // 2 + 1), np.arange(-N // 2 + 1, 0)))
    xi = np.fft.fft(rho) / ik
    xi[0] = 0
    ref = np.real(np.fft.ifft(xi))
    z = stockham(rho.astype(complex))
    b = np.imag(ik)
    zc = np.where(np.arange(N) == 0, 0, z.imag / b + 1j * z.real / b)
    got = np.real(stockham(zc)) / N
    assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()


def test_first_pass_stores_spread_over_the_banks():
    """Stride-8 stores of the first radix-8 pass (p = 1: point 8 i + r): with one pad slot per 8 points the 32 lanes of a warp
    cover all 32 four-byte banks four times over -- 512 B in the minimum of four wavefronts."""
    lane = np.arange(32)
    for r in range(8):
        slot = st_phys(8 * lane + r)          # 16-byte slots
        banks = (slot[:, None] * 4 + np.arange(4)[None, :]) % 32
        counts = np.bincount(banks.ravel(), minlength=32)
        assert counts.max() == 4 and counts.min() == 4
