set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "2d3v or 2d_stage or streamed" > gpurun_out/r2_10_tests.txt 2>&1; tail -6 gpurun_out/r2_10_tests.txt
for v in tiled ring43 ring23 ring42 ring82; do
  PICGOLF_2D_KERNEL=$v timeout 300 python bench.py --workload 2d3v --no-e2e --no-cpu --steps 20 --warmup 5 > gpurun_out/r2_10_2d3v_$v.json 2> gpurun_out/r2_10_2d3v_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2_10_2d3v_$v.json').read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['hbm_roofline_frac_step'], d['sorts_in_timed_region'], d['roofline']['stage_ms_per_step'])"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:particles_2d3v_ring -s 4 -c 1 -f -o gpurun_out/r2_10_ring43 python bench.py --workload 2d3v --no-e2e --no-cpu --steps 3 --warmup 3 > gpurun_out/r2_10_ncu.log 2>&1
tail -2 gpurun_out/r2_10_ncu.log | cut -c1-300
timeout 300 python tools/f_rows_timing.py > gpurun_out/r2_10_f_rows.jsonl 2> gpurun_out/r2_10_f_rows.err; cat gpurun_out/r2_10_f_rows.jsonl | cut -c1-600; tail -3 gpurun_out/r2_10_f_rows.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_10_bench.json 2> gpurun_out/r2_10_bench.err; tail -5 gpurun_out/r2_10_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_10_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'launches', d['gpu_launches'], 'roof', d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline']['stage_ms_per_step'])
print('e2e', json.dumps(d['e2e'])[:1200])
print('warm', json.dumps(d.get('warm_regime'))[:1500])
print('others', json.dumps(d['other_workloads'])[:2500])
print('clocks', d['clocks'])
print('cpu', json.dumps(d['cpu_baseline'])[:600])
PY
