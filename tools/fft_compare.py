#!/usr/bin/env python3
"""Timed comparison: libpicgolf's fused shared-memory FFT solve vs the same solve through cuFFT (torch.fft).

north_star keeps cuFFT "only as a timed comparison": cuFFT is never on the product path.  The library's solve
time comes from its own CUDA-event stage timers; the cuFFT arm does fft -> ./ik, xi[1]=0 -> ifft -> real with
torch complex128 tensors (3-4 launches), timed with CUDA events.  Run on the GPU box:  python tools/fft_compare.py
"""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particleincellcodegolf.jl_b200 as pg  # noqa: E402


def time_cuda(fn, iters=2000):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    dev = torch.device("cuda", 0)
    print("# fused shared-memory FFT solve (libpicgolf) vs cuFFT via torch.fft, complex128, B200")
    for N in (128, 4096):
        sim = pg.ngp_fourier(N=N, P=1 << 16, NT=8, W=float(N))
        sim.init_synthetic(seed=1)
        sim.step(8)
        sim.stage_timing(True); sim.stage_times(reset=True)
        sim.step(200)
        ours = sim.stage_times()["solve"] / 200 * 1e3
        rho = torch.rand(N, dtype=torch.float64, device=dev)
        kk = torch.tensor(np.concatenate([[1], np.arange(1, N // 2 + 1), np.arange(-N // 2 + 1, 0)]), dtype=torch.float64, device=dev)
        ik = (2j * math.pi) * kk

        def solve():
            xi = torch.fft.fft(rho) / ik
            xi[0] = 0
            return torch.fft.ifft(xi).real
        cu = time_cuda(solve)
        e_ours = pg.solve1d(rho.cpu().numpy())
        err = np.abs(e_ours - solve().cpu().numpy()).max() / np.abs(e_ours).max()
        print(f"1D N={N:5d}: solve1d_kernel {ours:7.1f} us/solve (one launch, fused ./ik, DC zero, norms, isapprox) | cuFFT path {cu:7.1f} us | max rel diff {err:.1e}")
    NX = NY = 256
    sim = pg.electrostatic_2d3v(NX=NX, NY=NY, P=1 << 18, T=8)
    sim.init_synthetic(seed=1, vth=sim.vth)
    sim.step(8)
    sim.stage_timing(True); sim.stage_times(reset=True)
    sim.step(200)
    ours = sim.stage_times()["solve"] / 200 * 1e3
    rho = torch.rand(NX, NY, dtype=torch.float64, device=dev)
    kx = 2 * math.pi * torch.tensor(np.concatenate([np.arange(0, NX // 2), np.arange(-NX // 2, 0)]), dtype=torch.float64, device=dev)
    k2 = kx[:, None] ** 2 + kx[None, :] ** 2
    k2[0, 0] = 1

    def solve2():
        phi = torch.fft.fft2(rho)
        phi[0, 0] = 0
        tmp = phi * (-1j / k2)
        return torch.fft.ifft2(tmp * kx[:, None]).real, torch.fft.ifft2(tmp * kx[None, :]).real
    cu = time_cuda(solve2, 500)
    print(f"2D {NX}x{NY}: solve2d_* (3 launches, 2 complex FFTs) {ours:7.1f} us/solve | cuFFT path (3 FFTs as in the reference) {cu:7.1f} us")


if __name__ == "__main__":
    main()
