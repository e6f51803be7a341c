set -x
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_51_gpu_tests.txt 2>&1; tail -4 gpurun_out/r2_51_gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_51_smoke.txt 2>&1; grep -c " ok" gpurun_out/r2_51_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_51_bench.json 2> gpurun_out/r2_51_bench.err; tail -3 gpurun_out/r2_51_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_51_bench_reference.json 2> gpurun_out/r2_51_bench_reference.err; tail -c 600 gpurun_out/r2_51_bench_reference.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_51_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches'], 'roof', d['roofline']['frac'], d['roofline']['launch_ms'], d['hbm_roofline_frac_step'], d['roofline']['stage_ms_per_step'])
print('e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['serial_set_step_get']['ms_per_step'])
print('warm', {k:(v['ms_per_step'], v['resorts'], v['resorts_fused_into_the_passes'], v['hbm_roofline_frac_step']) for k,v in d['warm_regime'].items() if isinstance(v,dict)})
print('others', {k:(v['ms_per_step'], v['hbm_roofline_frac_step']) for k,v in d['other_workloads'].items()})
print('cpu', d['cpu_baseline'])
print('clocks', d['clocks'])
PY
