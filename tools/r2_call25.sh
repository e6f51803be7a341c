set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "grid or solve1d or ngp" > gpurun_out/r2_25_tests.txt 2>&1; tail -4 gpurun_out/r2_25_tests.txt
python - <<'PY'
import numpy as np, torch
import particleincellcodegolf.jl_b200 as pg
for N in (1000, 4096):
    P = 1 << 26
    sim = pg.ngp_fourier(N=N, P=P, NT=64)
    sim.init_synthetic(seed=3)
    sim.step(4); sim.synchronize()
    st = torch.cuda.ExternalStream(sim.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); sim.step(20); e1.record(st); sim.synchronize(); torch.cuda.synchronize()
    print('N', N, 'P 2^26 ms/step', e0.elapsed_time(e1) / 20, 'G/s', P / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9)
    sim.close()
PY
