set -x
timeout 1500 python -m pytest tests -m gpu -q -x -k "deterministic or poly or streamed or config4 or c2" > gpurun_out/r2_gpu_tests_g.txt 2>&1; tail -25 gpurun_out/r2_gpu_tests_g.txt
python - <<'PY'
import time, numpy as np, torch
import particleincellcodegolf.jl_b200 as pg
for det in (0, 1):
    sim = pg.gaussian_fixed_point(N=4096, P=1 << 28, T=64, W=400.0, l=1e-8, deterministic=det)
    sim.init_synthetic(seed=1234)
    sim.step(5); sim.synchronize()
    st = torch.cuda.ExternalStream(sim.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); sim.step(20); e1.record(st); sim.synchronize(); torch.cuda.synchronize()
    print('det', det, 'path', sim.deposit_path, 'ms/step', e0.elapsed_time(e1) / 20, 'sorts', sim.sort_stats())
    sim.close()
PY
