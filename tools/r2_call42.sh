timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_42_gpu_tests.txt 2>&1; tail -4 gpurun_out/r2_42_gpu_tests.txt; grep -E "^E  " gpurun_out/r2_42_gpu_tests.txt | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_42_smoke.txt 2>&1; grep -c ok gpurun_out/r2_42_smoke.txt; tail -2 gpurun_out/r2_42_smoke.txt | cut -c1-200
timeout 600 python bench.py --workload ngp --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/r2_42_ngp.json 2> gpurun_out/r2_42_ngp.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_42_ngp.json').read().strip().splitlines()[-1])
print('ngp ms/step', d['ms_per_step'], d['roofline']['frac'], d.get('hbm_roofline_frac_step'), d['gpu_launches'])
PY
