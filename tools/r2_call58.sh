timeout 900 python -m pytest tests -m gpu -q -x -k "simpson or 1d2v or area or explicit or c2_fixed" 2>&1 | tail -2
timeout 600 python tools/f_rows_timing.py > gpurun_out/r2_58_f_rows_timing.jsonl 2>/dev/null; python - <<'PY'
import json
for l in open('gpurun_out/r2_58_f_rows_timing.jsonl'):
    try: d=json.loads(l)
    except: continue
    print(d['scheme'], round(d['ms_per_step'],3), 'ms/step', round(d['particle_steps_per_s']/1e9,2), 'G/s')
PY
