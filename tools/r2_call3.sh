set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_c.txt 2>&1; tail -15 gpurun_out/r2_gpu_tests_c.txt
V=particleincellcodegolf.jl_b200/lib/variants
for v in default mb3 mb5 mb6 t256mb2; do
  if [ $v = default ]; then unset PICGOLF_LIB; else export PICGOLF_LIB=$V/libpicgolf_$v.so; fi
  timeout 300 python bench.py --no-e2e --no-cpu --no-others --steps 20 --warmup 5 > gpurun_out/r2_c_$v.json 2> gpurun_out/r2_c_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_c_$v.json').read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['gpu_launches'], d['mean_sweeps_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['hbm_roofline_frac_step'], d['sorts_in_timed_region'], d['roofline']['stage_ms_per_step'])"
done
unset PICGOLF_LIB
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fp_pass_poly -s 12 -c 3 -f -o gpurun_out/r2_c_poly python bench.py --no-e2e --no-cpu --no-others --steps 3 --warmup 3 > gpurun_out/r2_c_ncu.log 2>&1
tail -2 gpurun_out/r2_c_ncu.log | cut -c1-300
