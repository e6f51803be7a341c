set -x
python - <<'PY'
import time, numpy as np, torch
import particleincellcodegolf.jl_b200 as pg
for det in (0, 1):
    sim = pg.gaussian_fixed_point(N=4096, P=1 << 28, T=64, W=400.0, l=1e-8, deterministic=det)
    sim.init_synthetic(seed=1234)
    sim.step(5); sim.synchronize()
    st = torch.cuda.ExternalStream(sim.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); sim.step(40); e1.record(st); sim.synchronize(); torch.cuda.synchronize()
    print('det', det, 'path', sim.deposit_path, 'ms/step', e0.elapsed_time(e1) / 40, 'sorts', sim.sort_stats())
    sim.close()
PY
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_h.txt 2>&1; tail -5 gpurun_out/r2_gpu_tests_h.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-warm --no-others > gpurun_out/r2_h_bench.json 2> gpurun_out/r2_h_bench.err; tail -3 gpurun_out/r2_h_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_h_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'roof', d['roofline']['frac'], d['roofline']['launch_ms'])
print('clocks', d['clocks'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['serial_set_step_get']['ms_per_step'])
PY
