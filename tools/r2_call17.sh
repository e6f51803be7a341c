set -x
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_17_gpu_tests.txt 2>&1; tail -5 gpurun_out/r2_17_gpu_tests.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_17_launches_2d3v.csv python bench.py --workload 2d3v --no-e2e --no-cpu --steps 20 --warmup 5 > gpurun_out/r2_17_ncu_2d3v.log 2>&1
grep -c sort_scatter gpurun_out/r2_17_launches_2d3v.csv; grep sort_scatter gpurun_out/r2_17_launches_2d3v.csv | cut -d, -f5,15- | head
timeout 300 python bench.py --workload 2d3v --no-e2e --no-cpu --steps 20 --warmup 5 > gpurun_out/r2_17_2d3v.json 2> gpurun_out/r2_17_2d3v.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_17_2d3v.json').read().strip().splitlines()[-1]); print('2d3v', round(d['ms_per_step'],3), round(d['roofline']['launch_ms'],3), round(d['roofline']['frac'],3), round(d['hbm_roofline_frac_step'],3), d['sorts_in_timed_region'])"
PICGOLF_ES_KERNEL=stream timeout 300 python tools/es_timing.py --shapes 1,12 --sort-every 0 --steps 48 > gpurun_out/r2_17_es_stream.jsonl 2> gpurun_out/r2_17_es_stream.err
cut -c1-330 gpurun_out/r2_17_es_stream.jsonl
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fp_pass_poly -s 13 -c 1 -f -o gpurun_out/r2_17_poly python bench.py --no-e2e --no-cpu --no-warm --no-others --steps 3 --warmup 3 > gpurun_out/r2_17_ncu_poly.log 2>&1
tail -1 gpurun_out/r2_17_ncu_poly.log | cut -c1-200
