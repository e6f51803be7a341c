set -x
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "grid or solve1d or c1_ngp or ngp_" > gpurun_out/r2_23_tests.txt 2>&1; tail -15 gpurun_out/r2_23_tests.txt
true
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "any_even_grid or not_a_power" > gpurun_out/r2_23_memcheck_dft.txt 2>&1; echo rc=$?; tail -4 gpurun_out/r2_23_memcheck_dft.txt
python - <<'PY'
import numpy as np, torch, time
import particleincellcodegolf.jl_b200 as pg
for N in (1000, 4096, 6000):
    P = 1 << 24
    sim = pg.ngp_fourier(N=N, P=P, NT=64)
    sim.init_synthetic(seed=3)
    sim.step(4); sim.synchronize()
    st = torch.cuda.ExternalStream(sim.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); sim.step(20); e1.record(st); sim.synchronize(); torch.cuda.synchronize()
    sim.stage_timing(True); sim.stage_times(reset=True); sim.step(8); t = sim.stage_times(reset=True)
    print('N', N, 'ms/step', e0.elapsed_time(e1) / 20, {k: v / 8 for k, v in t.items()})
    sim.close()
PY
