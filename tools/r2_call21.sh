set -x
timeout 300 python bench.py --workload ngp --no-e2e --no-cpu --steps 50 --warmup 5 > gpurun_out/r2_21_ngp.json 2> gpurun_out/r2_21_ngp.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_21_ngp.json').read().strip().splitlines()[-1]); print('ngp', d['ms_per_step'], d['gpu_launches'], d['roofline']['launch_ms'], d['roofline']['frac'], d['hbm_roofline_frac_step'], d['roofline']['stage_ms_per_step'])"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_21_smoke.txt 2>&1; tail -5 gpurun_out/r2_21_smoke.txt
