set -x
(which julia; ls baseline/_ref 2>&1; nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; free -g | head -2; nvidia-smi topo -m) > gpurun_out/r2_box_probe.txt 2>&1
PICGOLF_TEST_EXPERIMENTS=1 timeout 600 python -m pytest tests -m gpu -x -q -k experiment > gpurun_out/r2_experiment_tests.txt 2>&1
tail -5 gpurun_out/r2_experiment_tests.txt
timeout 300 python bench.py --workload 2d3v --no-e2e --no-cpu --steps 20 --warmup 5 > gpurun_out/r2_2d3v_base.json 2> gpurun_out/r2_2d3v_base.err
PICGOLF_2D_AGG=1 timeout 300 python bench.py --workload 2d3v --no-e2e --no-cpu --steps 20 --warmup 5 > gpurun_out/r2_2d3v_agg.json 2> gpurun_out/r2_2d3v_agg.err
timeout 300 python tools/es_timing.py --shapes 1,12 --sort-every 0 > gpurun_out/r2_es_base.jsonl 2>&1
PICGOLF_ES_AGG=1 timeout 300 python tools/es_timing.py --shapes 1,12 --sort-every 0 > gpurun_out/r2_es_agg.jsonl 2>&1
PICGOLF_2D_AGG=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:particles_2d3v_tiled -s 4 -c 1 -f -o gpurun_out/r2_2d_agg python bench.py --workload 2d3v --no-e2e --no-cpu --steps 3 --warmup 3 > gpurun_out/r2_ncu_agg.log 2>&1
cat gpurun_out/r2_box_probe.txt | head -30
for f in gpurun_out/r2_2d3v_base.json gpurun_out/r2_2d3v_agg.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['hbm_roofline_frac_step'], d['sorts_in_timed_region'])"; done
cat gpurun_out/r2_es_base.jsonl gpurun_out/r2_es_agg.jsonl | cut -c1-400
