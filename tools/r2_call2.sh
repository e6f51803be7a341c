set -x
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cond_while tools/micro/cond_while.cu && /tmp/cond_while > gpurun_out/r2_cond_while.txt 2>&1; cat gpurun_out/r2_cond_while.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_b.txt 2>&1; tail -8 gpurun_out/r2_gpu_tests_b.txt
timeout 300 python bench.py --no-e2e --no-cpu --no-others --steps 20 --warmup 5 > gpurun_out/r2_b_loop.json 2> gpurun_out/r2_b_loop.err; tail -3 gpurun_out/r2_b_loop.err
PICGOLF_LOOP=0 timeout 300 python bench.py --no-e2e --no-cpu --no-others --steps 20 --warmup 5 > gpurun_out/r2_b_fixed.json 2> gpurun_out/r2_b_fixed.err
for f in gpurun_out/r2_b_loop.json gpurun_out/r2_b_fixed.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['gpu_launches'], d['mean_sweeps_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['hbm_roofline_frac_step'], d['sorts_in_timed_region'], d['roofline']['stage_ms_per_step'])"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_b_launches_loop.csv python bench.py --no-e2e --no-cpu --no-others --steps 3 --warmup 3 > gpurun_out/r2_b_ncu.log 2>&1
tail -3 gpurun_out/r2_b_ncu.log
