set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_f.txt 2>&1; tail -5 gpurun_out/r2_gpu_tests_f.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_f_bench.json 2> gpurun_out/r2_f_bench.err; tail -5 gpurun_out/r2_f_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_f_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'roof', d['roofline']['frac'], d['roofline']['launch_ms'])
print('e2e', json.dumps(d['e2e'])[:1200])
print('warm', json.dumps(d['warm_regime'])[:1500])
print('others', json.dumps(d['other_workloads'])[:2500])
print('clocks', d['clocks'])
print('cpu', json.dumps(d['cpu_baseline'])[:600])
PY
