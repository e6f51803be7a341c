set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_d.txt 2>&1; tail -15 gpurun_out/r2_gpu_tests_d.txt
V=particleincellcodegolf.jl_b200/lib/variants
for v in default st3 st6 st8 t256st6; do
  if [ $v = default ]; then unset PICGOLF_LIB; else export PICGOLF_LIB=$V/libpicgolf_$v.so; fi
  timeout 300 python bench.py --no-e2e --no-cpu --no-others --steps 20 --warmup 5 > gpurun_out/r2_d_$v.json 2> gpurun_out/r2_d_$v.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_d_$v.json').read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['gpu_launches'], d['mean_sweeps_per_step'], d['roofline']['launch_ms'], d['roofline']['frac'], d['hbm_roofline_frac_step'], d['sorts_in_timed_region'], d['roofline']['stage_ms_per_step'])"
done
