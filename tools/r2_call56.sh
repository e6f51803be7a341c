timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_56_gpu_tests.txt 2>&1; tail -3 gpurun_out/r2_56_gpu_tests.txt; grep -E "^E  " gpurun_out/r2_56_gpu_tests.txt | head -5
timeout 600 python tools/f_rows_timing.py > gpurun_out/r2_56_f_rows_timing.jsonl 2>gpurun_out/r2_56_f_rows.err; python - <<'PY'
import json
for l in open('gpurun_out/r2_56_f_rows_timing.jsonl'):
    try: d=json.loads(l)
    except: continue
    print(d['scheme'], round(d['ms_per_step'],3), 'ms/step', round(d['particle_steps_per_s']/1e9,2), 'G/s')
PY
