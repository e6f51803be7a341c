timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_63_gpu_tests.txt 2>&1; tail -3 gpurun_out/r2_63_gpu_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_63_smoke.txt 2>&1; grep -c " ok" gpurun_out/r2_63_smoke.txt; grep "parity ledger" gpurun_out/r2_63_smoke.txt | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2_63_bench_default.json 2> gpurun_out/r2_63_bench_default.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_63_bench_default.json').read().strip().splitlines()[-1])
print('default flags: ms/step', d['ms_per_step'], 'value', d['value'], 'steps', d['steps'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
PY
