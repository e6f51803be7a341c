set -x
timeout 900 python -m pytest tests/test_zz_esfield_gpu.py -m gpu -x -q > gpurun_out/r2_14_es_tests.txt 2>&1; tail -3 gpurun_out/r2_14_es_tests.txt
for v in stream stream24_768 stream24_1024; do
  PICGOLF_ES_KERNEL=$v timeout 300 python tools/es_timing.py --shapes 1,12,13 --sort-every 0 --steps 48 > gpurun_out/r2_14_es_$v.jsonl 2> gpurun_out/r2_14_es_$v.err
  python - <<PY
import json
for l in open('gpurun_out/r2_14_es_$v.jsonl'):
    d=json.loads(l); print('$v', d['shape'], round(d['ms_per_step'],3), round(d['particle_steps_per_s']/1e9,2), round(d['hbm_frac_at_80B'],3), d['sorts_slow'], d['energy_drift'])
PY
  tail -2 gpurun_out/r2_14_es_$v.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_14_launches_2d3v.csv python bench.py --workload 2d3v --no-e2e --no-cpu --steps 20 --warmup 5 > gpurun_out/r2_14_ncu_2d3v.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_14_launches_gauss_fp.csv python bench.py --no-e2e --no-cpu --no-warm --no-others --steps 20 --warmup 5 > gpurun_out/r2_14_ncu_gauss.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_14_launches_ngp.csv python bench.py --workload ngp --no-e2e --no-cpu --steps 20 --warmup 5 > gpurun_out/r2_14_ncu_ngp.log 2>&1
ls -la gpurun_out/*.csv
