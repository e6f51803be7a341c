timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_43_launches_ngp.csv python bench.py --workload ngp --no-e2e --no-cpu --steps 8 --warmup 3 > gpurun_out/r2_43_ncu_ngp.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_43_launches_ngp.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[-9:]:
    print(r[4][:50].ljust(50), r[-1])
PY
for i in 1 2; do timeout 600 python bench.py --workload ngp --steps 40 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ngp ms/step', d['ms_per_step'], 'kernel', d['roofline']['launch_ms'], d['roofline']['frac'], d['gpu_launches'])"; done
