set -x
timeout 900 python -m pytest tests/test_zz_esfield_gpu.py -m gpu -x -q > gpurun_out/r2_13_es_tests.txt 2>&1; tail -6 gpurun_out/r2_13_es_tests.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_esfield_gpu.py -m gpu -x -q -k "tiled" > gpurun_out/r2_13_es_memcheck.txt 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/r2_13_es_memcheck.txt
for v in stream stream11 tiled; do
  PICGOLF_ES_KERNEL=$v timeout 300 python tools/es_timing.py --shapes 0,1,12,13,15 --sort-every 0 > gpurun_out/r2_13_es_$v.jsonl 2> gpurun_out/r2_13_es_$v.err
  python - <<PY
import json
for l in open('gpurun_out/r2_13_es_$v.jsonl'):
    d=json.loads(l); print('$v', d['shape'], round(d['ms_per_step'],3), round(d['particle_steps_per_s']/1e9,2), round(d['hbm_frac_at_80B'],3), d['sorts_slow'], d['energy_drift'])
PY
  tail -2 gpurun_out/r2_13_es_$v.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:es_particles_stream -s 6 -c 1 -f -o gpurun_out/r2_13_es_stream python tools/es_timing.py --shapes 12 --sort-every 0 --steps 4 > gpurun_out/r2_13_ncu.log 2>&1
tail -2 gpurun_out/r2_13_ncu.log | cut -c1-300
