set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_15_gpu_tests.txt 2>&1; tail -8 gpurun_out/r2_15_gpu_tests.txt
for v in stream stream11; do
  PICGOLF_ES_KERNEL=$v timeout 300 python tools/es_timing.py --shapes 0,1,12,13,15 --sort-every 0 --steps 48 > gpurun_out/r2_15_es_$v.jsonl 2> gpurun_out/r2_15_es_$v.err
  python - <<PY
import json
for l in open('gpurun_out/r2_15_es_$v.jsonl'):
    d=json.loads(l); print('$v', d['shape'], round(d['ms_per_step'],3), round(d['particle_steps_per_s']/1e9,2), round(d['hbm_frac_at_80B'],3), d['sorts_slow'], d['energy_drift'])
PY
  tail -2 gpurun_out/r2_15_es_$v.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:es_particles_stream -s 6 -c 1 -f -o gpurun_out/r2_15_es_stream python tools/es_timing.py --shapes 12 --sort-every 0 --steps 4 > gpurun_out/r2_15_ncu.log 2>&1
timeout 300 python bench.py --workload 2d3v --no-e2e --no-cpu --steps 20 --warmup 5 > gpurun_out/r2_15_2d3v.json 2> gpurun_out/r2_15_2d3v.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_15_2d3v.json').read().strip().splitlines()[-1]); print('2d3v', round(d['ms_per_step'],3), round(d['roofline']['launch_ms'],3), round(d['roofline']['frac'],3), round(d['hbm_roofline_frac_step'],3), d['sorts_in_timed_region'])"
