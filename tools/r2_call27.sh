set -x
timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "fused or poly or sorted or config4 or c2 or checkpoint or streamed" > gpurun_out/r2_27_tests.txt 2>&1; tail -15 gpurun_out/r2_27_tests.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-others --no-cpu > gpurun_out/r2_27_bench.json 2> gpurun_out/r2_27_bench.err; tail -3 gpurun_out/r2_27_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_27_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'roof', d['roofline']['frac'], d['roofline']['launch_ms'], d['hbm_roofline_frac_step'], d['roofline']['stage_ms_per_step'], 'sorts', d['sorts_in_timed_region'], d['sorts_fused_into_the_passes_since_start'])
print('e2e', d['e2e']['ms_per_step'])
print('warm', {k:(v['ms_per_step'], v['resorts'], v['resorts_fused_into_the_passes']) for k,v in d['warm_regime'].items() if isinstance(v,dict)})
PY
PICGOLF_FUSED_SORT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-others --no-cpu --no-e2e > gpurun_out/r2_27_bench_nofuse.json 2> gpurun_out/r2_27_bench_nofuse.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_27_bench_nofuse.json').read().strip().splitlines()[-1])
print('NOFUSE ms/step', d['ms_per_step'], 'warm', {k:(v['ms_per_step'], v['resorts']) for k,v in d['warm_regime'].items() if isinstance(v,dict)})
PY
